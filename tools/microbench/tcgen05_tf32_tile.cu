// Exploration for SURVEY 8f row 1 (batched policy inference = grouped GEMMs): one tcgen05 TF32 MMA
// tile, D[128 x N] = A[128 x K] * B[N x K]^T, operands written to shared memory by ordinary stores
// in the canonical K-major no-swizzle core-matrix layout, accumulator in TMEM, read back with
// tcgen05.ld.  Not part of the product; validates the descriptor encodings on this toolchain.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32;          // K in tf32 elements; one MMA consumes K = 8

__device__ __forceinline__ uint64_t make_desc(const void *smem, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem);
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);              // start address
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;   // leading dimension byte offset
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;   // stride dimension byte offset
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    return d;                                           // layout_type = 0: no swizzle
}

__global__ void __launch_bounds__(128) tile_kernel(const float *A, const float *B, float *D)
{
    // A: [K/4][M][4], B: [K/4][N][4] floats -> LBO = rows * 16 B (next 4 k), SBO = 128 B (next 8 rows)
    __shared__ __align__(128) float sA[K / 4][M][4];
    __shared__ __align__(128) float sB[K / 4][N][4];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < M * K; idx += 128) { const int m = idx / K, k = idx % K; sA[k / 4][m][k % 4] = A[idx]; }
    for (int idx = tid; idx < N * K; idx += 128) { const int n = idx / K, k = idx % K; sB[k / 4][n][k % 4] = B[idx]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(&tmem_base)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (2 at bits 7-9 / 10-12), K-major both,
        // N >> 3 at bits 17-22, M >> 4 at bits 24-28
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int k = 0; k < K / 8; ++k) {
            const uint64_t da = make_desc(&sA[2 * k][0][0], M * 16, 128);
            const uint64_t db = make_desc(&sB[2 * k][0][0], N * 16, 128);
            const uint32_t acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db),
                         "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
            (uint32_t)__cvta_generic_to_shared(&mbar)));
    }
    // everybody waits for the MMAs (phase 0 of the mbarrier)
    {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
        uint32_t done = 0;
        int spins = 0;
        while (!done && spins < (1 << 22)) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
            ++spins;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reads TMEM lanes 32 w .. 32 w + 31 (row m = lane), 64 columns, 8 at a time
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 8; ++q) D[(size_t)tid * N + c0 + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

int main()
{
    std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N, 0.f);
    for (int i = 0; i < M * K; ++i) A[i] = (float)((i * 7 + 3) % 17 - 8) * 0.25f;      // exactly representable in tf32
    for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5 + 1) % 13 - 6) * 0.5f;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; R[m * N + n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    tile_kernel<<<1, 128>>>(dA, dB, dD);
    cudaError_t err = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(err));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int i = 0; i < M * N; ++i) { const double e = fabs((double)D[i] - R[i]); if (!(e <= 1e-4)) ++bad; if (e > maxerr) maxerr = e; }
    printf("mismatches %d of %d, max |err| %.3g; D[0..3] = %g %g %g %g (ref %g %g %g %g)\n", bad, M * N, maxerr, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
    return bad != 0;
}
