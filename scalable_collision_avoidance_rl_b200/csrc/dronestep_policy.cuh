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
//   * one CTA = one agent x one tile of 128 environments (M = 128 = the TMEM lanes); the W2 chunks
//     arrive by TMA bulk copies (cp.async.bulk, double buffered, counted on an mbarrier);
//   * layer 1 (K = in_dim, tiny) on the CUDA cores, 32 hidden units at a time, written straight
//     into shared memory as the A operand of the next layer: K-major, no swizzle, 8 x 16-byte core
//     matrices (LBO = 128 rows x 16 B between k-groups of 4, SBO = 128 B between 8-row groups);
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

struct PolicyArgs {
    int E, n, in_dim, n_actions, real_bytes;
    unsigned seed_lo, seed_hi, stream;
    const void *z;            // Real [E][n][in_dim]
    const float *W1, *b1;     // [n][300][in_dim], [n][300]
    const float *W2p;         // [n][10 chunks][2 (hi, lo)][8 k-groups][304 rows][4]   operand layout
    const float *b2;          // [n][304] (zero padded)
    const float *W3, *b3;     // [n][A][304] (zero padded), [n][A]
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
    float A_hi[8][128][4], A_lo[8][128][4];                 // layer-1 chunk, operand layout (2 x 16 KB)
    float B[2][2][8][kPolNP][4];                             // two buffers of a W2 chunk, hi / lo (2 x 76 KB)
    float W1[kPolHidden][kPolMaxIn];                         // 19 KB
    float b1[kPolKP], b2[kPolNP];
    float b3[kPolMaxA];
    unsigned long long full[2], mma_done;                    // mbarriers: W2 chunk landed / MMAs of a chunk done
    uint32_t tmem_base;
};
constexpr uint32_t kPolChunkBytes = 2 * 8 * kPolNP * 4 * sizeof(float);   // 77,824 B, contiguous in global memory

__device__ __forceinline__ void pol_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 28)) __trap();                  // a lost arrival must fail loudly, not hang the device
    }
}
// TMA bulk copy (1-D, no tensor map): one thread moves a whole W2 chunk global -> shared; the bytes are
// counted on the mbarrier (complete_tx)
__device__ __forceinline__ void pol_bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    const uint32_t b = pol_smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pol_smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(b) : "memory");
}

template <typename Real>
__global__ void __launch_bounds__(128, 1) policy_kernel(const PolicyArgs a)
{
    using V2 = typename vec2_of<Real>::type;
    extern __shared__ __align__(128) unsigned char pol_smem_raw[];
    PolicySmem &sm = *reinterpret_cast<PolicySmem *>(pol_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int agent = blockIdx.y;
    const int e = blockIdx.x * 128 + tid;                    // this thread's environment = TMEM lane tid
    const bool live = e < a.E;
    const int in_dim = a.in_dim, A = a.n_actions;

    // ---- per-agent parameters -> shared memory
    for (int idx = tid; idx < kPolHidden * in_dim; idx += 128)
        sm.W1[idx / in_dim][idx % in_dim] = a.W1[(size_t)agent * kPolHidden * in_dim + idx];
    for (int idx = tid; idx < kPolKP; idx += 128) sm.b1[idx] = idx < kPolHidden ? a.b1[(size_t)agent * kPolHidden + idx] : 0.f;
    for (int idx = tid; idx < kPolNP; idx += 128) sm.b2[idx] = a.b2[(size_t)agent * kPolNP + idx];
    if (tid < A) sm.b3[tid] = a.b3[(size_t)agent * A + tid];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pol_smem_u32(&sm.full[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pol_smem_u32(&sm.full[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pol_smem_u32(&sm.mma_done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pol_smem_u32(&sm.tmem_base)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // observation of this thread's (environment, agent): float32(z) as the reference casts it (utils.py:305)
    float zin[kPolMaxIn];
#pragma unroll
    for (int d = 0; d < kPolMaxIn; ++d)
        zin[d] = (live && d < in_dim) ? (float)reinterpret_cast<const Real *>(a.z)[((size_t)e * a.n + agent) * in_dim + d] : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm.tmem_base;

    // ---- layers 1 + 2, 32 hidden units of layer 1 (= 32 of K) at a time.  The W2 chunk of step kc + 2
    // is in flight (TMA bulk copy into the buffer the MMAs of step kc have just released) while step
    // kc + 1 computes layer 1 and runs its MMAs.
    const float *W2p = a.W2p + (size_t)agent * (kPolKP / kPolChunk) * (kPolChunkBytes / sizeof(float));
    constexpr int NCH = kPolKP / kPolChunk;
    if (tid == 0) {
        pol_bulk_load(&sm.B[0][0][0][0][0], W2p, kPolChunkBytes, &sm.full[0]);
        pol_bulk_load(&sm.B[1][0][0][0][0], W2p + kPolChunkBytes / sizeof(float), kPolChunkBytes, &sm.full[1]);
    }
    for (int kc = 0; kc < NCH; ++kc) {
        const int bsel = kc & 1;
        // layer 1 for this thread's environment: units 32 kc .. 32 kc + 31 (utils.py:289-290), split hi / lo
        // (A_hi / A_lo are free: the MMAs of step kc - 1 were waited for at the end of that step)
#pragma unroll
        for (int kg = 0; kg < 8; ++kg) {
            float hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = kc * kPolChunk + kg * 4 + q;
                float h = 0.f;
                if (j < kPolHidden) {
                    h = sm.b1[j];
#pragma unroll
                    for (int d = 0; d < kPolMaxIn; ++d)
                        if (d < in_dim) h = fmaf(sm.W1[j][d], zin[d], h);
                    h = fmaxf(h, 0.f);
                }
                hi[q] = __uint_as_float(__float_as_uint(h) & 0xffffe000u);       // what kind::tf32 reads
                lo[q] = h - hi[q];                                               // exact
            }
            *reinterpret_cast<float4 *>(&sm.A_hi[kg][tid][0]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(&sm.A_lo[kg][tid][0]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic stores -> async proxy
        __syncthreads();
        if (tid == 0) {
            pol_wait(pol_smem_u32(&sm.full[bsel]), (uint32_t)(kc >> 1) & 1u);    // W2 chunk kc has landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {                                     // one MMA = K of 8 = 2 k-groups
                const uint64_t a_hi = pol_desc(&sm.A_hi[2 * ks][0][0], 128 * 16, 128);
                const uint64_t a_lo = pol_desc(&sm.A_lo[2 * ks][0][0], 128 * 16, 128);
#pragma unroll
                for (int half = 0; half < 2; ++half) {                           // N = 160 + 144
                    const int row0 = half ? 160 : 0, N = half ? 144 : 160;
                    const uint64_t b_hi = pol_desc(&sm.B[bsel][0][2 * ks][row0][0], kPolNP * 16, 128);
                    const uint64_t b_lo = pol_desc(&sm.B[bsel][1][2 * ks][row0][0], kPolNP * 16, 128);
                    const uint32_t idesc = pol_idesc(N), d = tmem + (uint32_t)row0;
                    pol_mma(d, a_hi, b_hi, idesc, (kc | ks) != 0);
                    pol_mma(d, a_hi, b_lo, idesc, 1);
                    pol_mma(d, a_lo, b_hi, idesc, 1);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                pol_smem_u32(&sm.mma_done)));
        }
        // A_hi / A_lo and this W2 buffer may be overwritten (and, after the last step, D read) once the
        // MMAs are done
        pol_wait(pol_smem_u32(&sm.mma_done), (uint32_t)kc & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0 && kc + 2 < NCH)
            pol_bulk_load(&sm.B[bsel][0][0][0][0], W2p + (size_t)(kc + 2) * (kPolChunkBytes / sizeof(float)), kPolChunkBytes,
                          &sm.full[bsel]);
    }
    // W3 of this agent -> the (now free) first W2 buffer, as [A][304]
    float (*sW3)[kPolNP] = reinterpret_cast<float (*)[kPolNP]>(&sm.B[0][0][0][0][0]);
    for (int idx = tid; idx < A * kPolNP; idx += 128) sW3[idx / kPolNP][idx % kPolNP] = a.W3[(size_t)agent * A * kPolNP + idx];
    __syncthreads();

    // ---- epilogue: row e of D -> +b2, ReLU (utils.py:293-294) -> layer 3 on the CUDA cores (:297)
    float logit[kPolMaxA];
#pragma unroll
    for (int q = 0; q < kPolMaxA; ++q) logit[q] = (q < A) ? sm.b3[q] : -INFINITY;
    for (int c0 = 0; c0 < kPolNP; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int cidx = c0 + q;
            const float h2 = fmaxf(__uint_as_float(v[q]) + sm.b2[cidx], 0.f);     // padded columns: 0 + 0
#pragma unroll
            for (int aa = 0; aa < kPolMaxA; ++aa)
                if (aa < A) logit[aa] = fmaf(sW3[aa][cidx], h2, logit[aa]);
        }
    }
    // softmax over the actions (utils.py:298), index by inverse CDF of a Philox uniform (:307)
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < kPolMaxA; ++q) mx = fmaxf(mx, logit[q]);
    float pr[kPolMaxA], sum = 0.f;
#pragma unroll
    for (int q = 0; q < kPolMaxA; ++q) { pr[q] = (q < A) ? expf(logit[q] - mx) : 0.f; sum += pr[q]; }
    if (live) {
        unsigned rnd[4];
        philox4x32_10((unsigned)e, (unsigned)agent, a.stream, 1u, a.seed_lo, a.seed_hi, rnd);
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
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}
#endif  // __CUDACC__

}  // namespace ds
