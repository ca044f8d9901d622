// dronestep_policy.cuh -- batched policy inference on the 5th-generation tensor cores
// (SURVEY.md section 8f row 1).
//
// The reference's actors are per-agent MLPs, DiscreteSoftmaxNN (utils.py:255-318):
//   probs = softmax(W3 relu(W2 relu(W1 z + b1) + b2) + b3),  z = float32(z_state) (6 or 15 inputs),
//   300 -> 300 hidden units, A <= 16 actions; sample_action draws an index from probs and
//   returns action_list[index] (utils.py:304-309).  SA2CAgents.forward / TrainedAgent.forward call
//   it once per agent per step on the host (SAC_agents.py:60-82,170-180).
// Batched over E environments the 300 x 300 layer is a grouped GEMM -- n groups (one network per
// agent), M = E rows each -- and belongs on tcgen05:
//   * a tile = one agent x 128 environments (M = 128 = the TMEM lanes); persistent CTAs (one per SM)
//     walk contiguous ranges of tiles; the W2 chunks
//     arrive by TMA bulk copies (cp.async.bulk, double buffered, counted on an mbarrier);
//     warps 0-3 own the 128 rows, one thread of warp 4 issues copies and MMAs (no CTA barrier
//     in the main loop: mbarrier handshakes both ways);
//   * layer 1 (K = in_dim, tiny) on the CUDA cores, 16 hidden units at a time, written straight
//     into shared memory as the A operand of the next layer: K-major, no swizzle, 8 x 16-byte core
//     matrices (LBO = 128 rows x 16 B between k-groups of 4, SBO = 128 B between 8-row groups);
//     two such half-chunks alternate, so the MMAs of one run while the next is computed;
//   * layer 2 as tcgen05.mma.kind::tf32 with the accumulator D[128 x 304] in TMEM; fp32 parity is
//     kept by the 3xTF32 split: x = hi + lo with hi = the 19 bits the tensor core reads,
//     D += A_hi B_hi + A_hi B_lo + A_lo B_hi (relative error ~2^-20 instead of 2^-10);  W2 is
//     split and packed into the operand layout once on the host (ds_policy_create);
//   * epilogue: each thread owns one environment row (tcgen05.ld 32x32b), adds b2, ReLU, and folds
//     the 300 values into its <= 16 logits on the CUDA cores (layer 3), then softmax, and the
//     action index from a Philox uniform by inverse CDF.
// One elected thread issues the MMAs; completion comes back through tcgen05.commit -> mbarrier.
#pragma once
#include "dronestep_kernels.cuh"

namespace ds {

constexpr int kPolHidden = 300;       // DiscreteSoftmaxNN: Ls = hidden_1 = 300 (utils.py:272-273)
constexpr int kPolNP = 304;           // N of layer 2 padded to a multiple of 16 (two MMAs: 160 + 144)
constexpr int kPolKP = 320;           // K of layer 2 padded to whole chunks of 32
constexpr int kPolChunk = 32;         // hidden units of layer 1 produced per staging round
constexpr int kPolMaxIn = 16, kPolMaxA = 16;

// per-agent parameter block ("head"), packed by ds_policy_create so that one bulk copy brings it in
constexpr int kPolHeadW1 = 0;                                  // W1 transposed [16 inputs][320 units], zero padded
constexpr int kPolHeadB1 = kPolMaxIn * kPolKP;                 // b1 [320], zero padded
constexpr int kPolHeadB2 = kPolHeadB1 + kPolKP;                // b2 [304], zero padded
constexpr int kPolHeadB3 = kPolHeadB2 + kPolNP;                // b3 [16], -inf padded
constexpr int kPolHeadFloats = kPolHeadB3 + kPolMaxA;          // 5760 floats = 23,040 B

struct PolicyArgs {
    int E, n, in_dim, n_actions, real_bytes;
    unsigned seed_lo, seed_hi, stream;
    int tiles_per_agent;      // ceil(E / 128); tile t = (agent t / tiles_per_agent, environments 128 (t % tiles_per_agent) ..)
    const unsigned long long *seed_dev;   // when non-null the seed is read from device memory (graph replays)
    const void *z;            // Real [E][n][in_dim]
    const float *head;        // [n][kPolHeadFloats]
    const float *W2p;         // [n][10 chunks][2 (hi, lo)][8 k-groups][304 rows][4]   operand layout
    const float *W3t;         // [n][304][16]: W3 transposed, zero padded
    const void *atable;       // Real [A][2]
    void *act;                // Real [E][n][2] out
    uint8_t *aidx;            // [E][n] out (may be null)
    float *probs;             // [E][n][A] out (may be null)
};

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t pol_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, no swizzle (layout_type 0), sm_100 version bit
__device__ __forceinline__ uint64_t pol_desc(const void *smem, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = (uint64_t)((pol_smem_u32(smem) >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D f32, A / B tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t pol_idesc(int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void pol_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(accumulate));
}

struct PolicySmem {
    float A_hi[2][4][128][4], A_lo[2][4][128][4];           // two half-chunks (16 units) of layer 1, operand layout (32 KB)
    float B[2][2][8][kPolNP][4];                             // two buffers of a W2 chunk, hi / lo (2 x 76 KB)
    float W1t[kPolMaxIn][kPolKP];                            // the head block, in ds_policy_create's order (22.5 KB)
    float b1[kPolKP], b2[kPolNP], b3[kPolMaxA];
    float W3t[kPolNP][kPolMaxA];                             // 19 KB
    unsigned long long full[2], mma_done[2], a_ready[2], params, rows_done;   // mbarriers: W2 chunk landed / MMAs of a
                                                             // half-chunk done / A half written by the 128 rows / head + W3t
                                                             // landed / the rows have left a tile (parameters may change)
    uint32_t tmem_base;
};
static_assert(offsetof(PolicySmem, b1) - offsetof(PolicySmem, W1t) == kPolHeadB1 * sizeof(float) &&
              offsetof(PolicySmem, b2) - offsetof(PolicySmem, W1t) == kPolHeadB2 * sizeof(float) &&
              offsetof(PolicySmem, b3) - offsetof(PolicySmem, W1t) == kPolHeadB3 * sizeof(float), "head block layout");
constexpr uint32_t kPolChunkBytes = 2 * 8 * kPolNP * 4 * sizeof(float);   // 77,824 B, contiguous in global memory
constexpr uint32_t kPolHeadBytes = kPolHeadFloats * sizeof(float), kPolW3tBytes = kPolNP * kPolMaxA * sizeof(float);

__device__ __forceinline__ void pol_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 28)) __trap();                  // a lost arrival must fail loudly, not hang the device
    }
}
// TMA bulk copies (1-D, no tensor map): one thread moves a whole block global -> shared; the bytes are
// counted on the mbarrier (expect_tx / complete_tx)
__device__ __forceinline__ void pol_expect(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pol_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pol_bulk_copy(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pol_smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(pol_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pol_bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    pol_expect(bar, bytes);
    pol_bulk_copy(dst, src, bytes, bar);
}
// 16 consecutive accumulator columns of this thread's TMEM lane (asynchronous: pol_ld_wait before use)
__device__ __forceinline__ void pol_ld16(uint32_t (&v)[16], uint32_t taddr)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void pol_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// +b2, ReLU (utils.py:293-294) and layer 3 (:297) for 16 columns of the row: two actions per FFMA2
__device__ __forceinline__ void pol_layer3_block(float2 (&lg)[kPolMaxA / 2], const uint32_t (&v)[16], int c0,
                                                 const PolicySmem &sm, bool wide)
{
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int cidx = c0 + q;
        const float h2 = fmaxf(__uint_as_float(v[q]) + sm.b2[cidx], 0.f);         // padded columns: 0 + 0
        const float2 hh = make_float2(h2, h2);
        const float4 w0 = *reinterpret_cast<const float4 *>(&sm.W3t[cidx][0]);
        const float4 w1 = *reinterpret_cast<const float4 *>(&sm.W3t[cidx][4]);
        lg[0] = __ffma2_rn(make_float2(w0.x, w0.y), hh, lg[0]);
        lg[1] = __ffma2_rn(make_float2(w0.z, w0.w), hh, lg[1]);
        lg[2] = __ffma2_rn(make_float2(w1.x, w1.y), hh, lg[2]);
        lg[3] = __ffma2_rn(make_float2(w1.z, w1.w), hh, lg[3]);
        if (wide) {
            const float4 w2 = *reinterpret_cast<const float4 *>(&sm.W3t[cidx][8]);
            const float4 w3 = *reinterpret_cast<const float4 *>(&sm.W3t[cidx][12]);
            lg[4] = __ffma2_rn(make_float2(w2.x, w2.y), hh, lg[4]);
            lg[5] = __ffma2_rn(make_float2(w2.z, w2.w), hh, lg[5]);
            lg[6] = __ffma2_rn(make_float2(w3.x, w3.y), hh, lg[6]);
            lg[7] = __ffma2_rn(make_float2(w3.z, w3.w), hh, lg[7]);
        }
    }
}

// IN = in_dim when it is one of the reference's two observation widths (6: simplify_zstate, 15: full,
// k = 2), else 0 (any in_dim <= 16, guarded loop).  Layers 1 and 3 run on the packed-f32 pipe
// (FFMA2: two hidden units / two actions per instruction, operands read as float4).
// Warp roles: warps 0-3 = one thread per environment row (layer 1, epilogue); warp 4 = one elected
// thread that issues the bulk copies and the MMAs.  The main loop has no CTA barrier: rows -> issuer
// through a_ready (128 arrivals), issuer -> rows through tcgen05.commit on mma_done.
constexpr int kPolThreads = 160;
// PERSISTENT: the grid has one CTA per SM (or per tile, if fewer); a CTA walks a contiguous range of
// tiles, so that consecutive tiles mostly belong to the same agent: the parameter block stays in
// shared memory, the W2 ring simply continues (chunks 0 and 1 of the next tile arrive during the
// epilogue of the current one), TMEM and the mbarriers are set up once.  Step and chunk counters run
// across tiles; every mbarrier parity is derived from them.
template <typename Real, int IN>
__global__ void __launch_bounds__(kPolThreads, 1) policy_kernel(const PolicyArgs a)
{
    using V2 = typename vec2_of<Real>::type;
    extern __shared__ __align__(128) unsigned char pol_smem_raw[];
    PolicySmem &sm = *reinterpret_cast<PolicySmem *>(pol_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool row_thread = tid < 128, issuer = tid == 128;
    const int in_dim = a.in_dim, A = a.n_actions, TPA = a.tiles_per_agent;
    const long long total = (long long)a.n * TPA;
    const int t_begin = (int)(total * blockIdx.x / gridDim.x), t_end = (int)(total * (blockIdx.x + 1) / gridDim.x);
    constexpr int NST = 2 * (kPolKP / kPolChunk);            // 20 half-chunks of 16 hidden units (= 16 of K) per tile
    constexpr int NCH = kPolKP / kPolChunk;                  // 10 W2 chunks per tile
    constexpr size_t kChunkFloats = kPolChunkBytes / sizeof(float);
    auto w2_of = [&](int agent) { return a.W2p + (size_t)agent * NCH * kChunkFloats; };
    auto load_params = [&](int agent) {
        pol_expect(&sm.params, kPolHeadBytes + kPolW3tBytes);
        pol_bulk_copy(&sm.W1t[0][0], a.head + (size_t)agent * kPolHeadFloats, kPolHeadBytes, &sm.params);
        pol_bulk_copy(&sm.W3t[0][0], a.W3t + (size_t)agent * kPolNP * kPolMaxA, kPolW3tBytes, &sm.params);
    };
    // observation of a row thread's (environment, agent) of tile t: float32(z) as the reference casts it (utils.py:305)
    auto load_z = [&](float (&zz)[kPolMaxIn], int t) {
        const int agent = t / TPA, e = (t - agent * TPA) * 128 + tid;
        const bool ok = row_thread && t < t_end && e < a.E;
#pragma unroll
        for (int d = 0; d < kPolMaxIn; ++d)
            zz[d] = (ok && d < in_dim) ? (float)reinterpret_cast<const Real *>(a.z)[((size_t)e * a.n + agent) * in_dim + d] : 0.f;
    };

    // ---- once per CTA: mbarriers, TMEM, parameter block and W2 chunks 0 and 1 of the first tile
    if (issuer) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pol_smem_u32(&sm.full[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pol_smem_u32(&sm.full[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pol_smem_u32(&sm.mma_done[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pol_smem_u32(&sm.mma_done[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(pol_smem_u32(&sm.a_ready[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(pol_smem_u32(&sm.a_ready[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pol_smem_u32(&sm.params)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(pol_smem_u32(&sm.rows_done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int agent0 = t_begin / TPA;
        load_params(agent0);
        pol_bulk_load(&sm.B[0][0][0][0][0], w2_of(agent0), kPolChunkBytes, &sm.full[0]);
        pol_bulk_load(&sm.B[1][0][0][0][0], w2_of(agent0) + kChunkFloats, kPolChunkBytes, &sm.full[1]);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pol_smem_u32(&sm.tmem_base)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    float zin[kPolMaxIn], znext[kPolMaxIn];
    load_z(znext, t_begin);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm.tmem_base;

    if (issuer) {
        // ---- layer 2: MMAs of step s from A half (s & 1) as soon as the rows have written it.  W2 arrives in
        // chunks of 32 (two steps); the next chunk of the ring is copied into the buffer that the MMAs of the
        // previous step were the last to read.
        unsigned gs = 0;                                                     // steps issued so far (all tiles)
        for (int t = t_begin; t < t_end; ++t) {
            const int agent = t / TPA;
            const float *W2p = w2_of(agent);
            if (t > t_begin && agent != (t - 1) / TPA) {
                // new agent: its parameter block may land once every row has left the previous tile
                pol_wait(pol_smem_u32(&sm.rows_done), (uint32_t)(t - t_begin - 1) & 1u);
                load_params(agent);
            }
            for (int s_ = 0; s_ < NST; ++s_, ++gs) {
                const int ab = s_ & 1, kc = s_ >> 1, bsel = kc & 1;
                const unsigned g = (unsigned)(t - t_begin) * NCH + (unsigned)kc;    // chunk number in the ring
                if (ab == 0) pol_wait(pol_smem_u32(&sm.full[bsel]), (g >> 1) & 1u);  // W2 chunk has landed
                pol_wait(pol_smem_u32(&sm.a_ready[ab]), (gs >> 1) & 1u);             // A half written, fenced
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {                                     // one MMA = K of 8 = 2 k-groups
                    const uint64_t a_hi = pol_desc(&sm.A_hi[ab][2 * ks][0][0], 128 * 16, 128);
                    const uint64_t a_lo = pol_desc(&sm.A_lo[ab][2 * ks][0][0], 128 * 16, 128);
                    const int kgb = 4 * ab + 2 * ks;                                 // k-group inside the W2 chunk
#pragma unroll
                    for (int half = 0; half < 2; ++half) {                           // N = 160 + 144
                        const int row0 = half ? 160 : 0, N = half ? 144 : 160;
                        const uint64_t b_hi = pol_desc(&sm.B[bsel][0][kgb][row0][0], kPolNP * 16, 128);
                        const uint64_t b_lo = pol_desc(&sm.B[bsel][1][kgb][row0][0], kPolNP * 16, 128);
                        const uint32_t idesc = pol_idesc(N), d = tmem + (uint32_t)row0;
                        pol_mma(d, a_hi, b_hi, idesc, (s_ | ks) != 0);
                        pol_mma(d, a_hi, b_lo, idesc, 1);
                        pol_mma(d, a_lo, b_hi, idesc, 1);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                    pol_smem_u32(&sm.mma_done[ab])));
                if (ab == 0 && gs >= 2) {
                    const float *next = (kc + 1 < NCH) ? W2p + (size_t)(kc + 1) * kChunkFloats
                                                       : (t + 1 < t_end ? w2_of((t + 1) / TPA) : nullptr);
                    if (next) {
                        pol_wait(pol_smem_u32(&sm.mma_done[1]), ((gs - 1) >> 1) & 1u);
                        pol_bulk_load(&sm.B[bsel ^ 1][0][0][0][0], next, kPolChunkBytes, &sm.full[bsel ^ 1]);
                    }
                }
            }
        }
    } else if (row_thread) {
        unsigned gs = 0, nparam = 0;
        int prev_agent = -1;
        for (int t = t_begin; t < t_end; ++t) {
        const int agent = t / TPA;
        const int e = (t - agent * TPA) * 128 + tid;             // this thread's environment = TMEM lane tid
        const bool live = e < a.E;
        if (agent != prev_agent) {                               // the parameter block of this agent has landed
            pol_wait(pol_smem_u32(&sm.params), nparam & 1u);
            ++nparam; prev_agent = agent;
        }
#pragma unroll
        for (int d = 0; d < kPolMaxIn; ++d) zin[d] = znext[d];
        // ---- layer 1 for this thread's environment, 16 units per step (utils.py:289-290), split hi / lo
        for (int s_ = 0; s_ < NST; ++s_, ++gs) {
            const int ab = s_ & 1;
            // A half ab was last read by the MMAs of step gs - 2
            if (gs >= 2) pol_wait(pol_smem_u32(&sm.mma_done[ab]), ((gs >> 1) - 1) & 1u);
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
                const int j0 = s_ * 16 + kg * 4;                                 // units j0 .. j0 + 3 (padding: 0 weights)
                const float4 bb = *reinterpret_cast<const float4 *>(&sm.b1[j0]);
                float2 h01 = make_float2(bb.x, bb.y), h23 = make_float2(bb.z, bb.w);
#pragma unroll
                for (int d = 0; d < (IN > 0 ? IN : kPolMaxIn); ++d) {
                    if (IN > 0 || d < in_dim) {
                        const float4 w = *reinterpret_cast<const float4 *>(&sm.W1t[d][j0]);
                        const float2 zz = make_float2(zin[d], zin[d]);
                        h01 = __ffma2_rn(make_float2(w.x, w.y), zz, h01);
                        h23 = __ffma2_rn(make_float2(w.z, w.w), zz, h23);
                    }
                }
                const float h[4] = {fmaxf(h01.x, 0.f), fmaxf(h01.y, 0.f), fmaxf(h23.x, 0.f), fmaxf(h23.y, 0.f)};
                float hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    hi[q] = __uint_as_float(__float_as_uint(h[q]) & 0xffffe000u);    // what kind::tf32 reads
                    lo[q] = h[q] - hi[q];                                            // exact
                }
                *reinterpret_cast<float4 *>(&sm.A_hi[ab][kg][tid][0]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(&sm.A_lo[ab][kg][tid][0]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic stores -> async proxy
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");     // (the previous tile's TMEM reads)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pol_smem_u32(&sm.a_ready[ab])) : "memory");
        }
        load_z(znext, t + 1);                                    // the next tile's observation, in flight during the epilogue
        // D may be read once the MMAs of the last step (and with them all earlier ones) are done
        pol_wait(pol_smem_u32(&sm.mma_done[1]), ((gs - 1) >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: row e of D -> +b2, ReLU -> layer 3 on the CUDA cores; the tcgen05.ld of the next
    // 16 columns is in flight while the current 16 are folded into the logits
    float2 lg[kPolMaxA / 2];                                                   // logits, two actions per register pair
#pragma unroll
    for (int q = 0; q < kPolMaxA / 2; ++q) lg[q] = make_float2(sm.b3[2 * q], sm.b3[2 * q + 1]);   // -inf beyond A
    const bool wide = A > 8;                                                   // uniform: second half of the actions
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t va[16], vb[16];
    pol_ld16(va, trow);
    for (int c0 = 0; c0 < kPolNP; c0 += 32) {                                  // 19 blocks of 16: 9 pairs + 1
        pol_ld_wait();
        if (c0 + 16 < kPolNP) pol_ld16(vb, trow + (uint32_t)(c0 + 16));
        pol_layer3_block(lg, va, c0, sm, wide);
        if (c0 + 16 < kPolNP) {
            pol_ld_wait();
            if (c0 + 32 < kPolNP) pol_ld16(va, trow + (uint32_t)(c0 + 32));
            pol_layer3_block(lg, vb, c0 + 16, sm, wide);
        }
    }
    float logit[kPolMaxA];
#pragma unroll
    for (int q = 0; q < kPolMaxA / 2; ++q) { logit[2 * q] = lg[q].x; logit[2 * q + 1] = lg[q].y; }
    // softmax over the actions (utils.py:298), index by inverse CDF of a Philox uniform (:307)
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < kPolMaxA; ++q) mx = fmaxf(mx, logit[q]);
    float pr[kPolMaxA], sum = 0.f;
#pragma unroll
    for (int q = 0; q < kPolMaxA; ++q) { pr[q] = (q < A) ? expf(logit[q] - mx) : 0.f; sum += pr[q]; }
    if (live) {
        unsigned rnd[4];
        const unsigned long long seed = a.seed_dev ? *a.seed_dev : (((unsigned long long)a.seed_hi << 32) | a.seed_lo);
        philox4x32_10((unsigned)e, (unsigned)agent, a.stream, 1u, (unsigned)seed, (unsigned)(seed >> 32), rnd);
        const float u = (float)(rnd[0] >> 8) * 5.9604644775390625e-08f;       // [0, 1), 24 bits
        float cdf = 0.f;
        int pick = A - 1;
        bool found = false;
        const size_t ga = (size_t)e * a.n + agent;
#pragma unroll
        for (int q = 0; q < kPolMaxA; ++q) {
            if (q < A) {
                const float p = pr[q] / sum;
                if (a.probs) a.probs[ga * A + q] = p;
                cdf += p;
                if (!found && u < cdf) { pick = q; found = true; }
            }
        }
        reinterpret_cast<V2 *>(a.act)[ga] = reinterpret_cast<const V2 *>(a.atable)[pick];   // action_list[arg] (:309)
        if (a.aidx) a.aidx[ga] = (uint8_t)pick;
    }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pol_smem_u32(&sm.rows_done)) : "memory");
        }   // tiles
    }   // row threads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
#endif  // __CUDACC__

}  // namespace ds
