// Microbenchmark (tuning aid, not part of the product): FP64 pipe throughput / latency on the SM,
// and the cost of the correctly rounded sqrt / div and of log, per SM, as a function of resident warps.
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>

template <int MODE>
__global__ void k(double *out, int iters, long long *cycles)
{
    double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3;
    double a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
    const double b = 1.0000001, c = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {           // 8 independent DFMA chains
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        } else if (MODE == 1) {    // one dependent DFMA chain
            a0 = fma(a0, b, c); a0 = fma(a0, b, c); a0 = fma(a0, b, c); a0 = fma(a0, b, c);
            a0 = fma(a0, b, c); a0 = fma(a0, b, c); a0 = fma(a0, b, c); a0 = fma(a0, b, c);
        } else if (MODE == 2) {    // sqrt_rn x8 independent
            a0 = __dsqrt_rn(a0 + 1.0); a1 = __dsqrt_rn(a1 + 1.0); a2 = __dsqrt_rn(a2 + 1.0); a3 = __dsqrt_rn(a3 + 1.0);
            a4 = __dsqrt_rn(a4 + 1.0); a5 = __dsqrt_rn(a5 + 1.0); a6 = __dsqrt_rn(a6 + 1.0); a7 = __dsqrt_rn(a7 + 1.0);
        } else if (MODE == 3) {    // div_rn x8
            a0 = __ddiv_rn(b, a0 + 1.0); a1 = __ddiv_rn(b, a1 + 1.0); a2 = __ddiv_rn(b, a2 + 1.0); a3 = __ddiv_rn(b, a3 + 1.0);
            a4 = __ddiv_rn(b, a4 + 1.0); a5 = __ddiv_rn(b, a5 + 1.0); a6 = __ddiv_rn(b, a6 + 1.0); a7 = __ddiv_rn(b, a7 + 1.0);
        } else if (MODE == 4) {    // log x8
            a0 = log(a0 + 1.5); a1 = log(a1 + 1.5); a2 = log(a2 + 1.5); a3 = log(a3 + 1.5);
            a4 = log(a4 + 1.5); a5 = log(a5 + 1.5); a6 = log(a6 + 1.5); a7 = log(a7 + 1.5);
        } else if (MODE == 5) {    // 8 independent DADD
            a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
        } else if (MODE == 6) {    // DSETP + SEL mix: compare chains
            a0 = (a0 < a1) ? a0 + c : a1; a2 = (a2 < a3) ? a2 + c : a3; a4 = (a4 < a5) ? a4 + c : a5; a6 = (a6 < a7) ? a6 + c : a7;
            a1 = (a1 < a0) ? a1 + c : a0; a3 = (a3 < a2) ? a3 + c : a2; a5 = (a5 < a4) ? a5 + c : a4; a7 = (a7 < a6) ? a7 + c : a6;
        } else if (MODE == 7) {    // 4 DFMA + 4 independent integer ops (co-issue?)
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            unsigned x = __double2loint(a4), y = __double2loint(a5);
            x = x * 1664525u + 1013904223u; y = (y ^ x) + (x >> 3);
            a4 = __hiloint2double(__double2hiint(a4), x); a5 = __hiloint2double(__double2hiint(a5), y);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ops_per_iter, double *out, long long *cyc, int sms)
{
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        const int threads = 32 * (warps < 32 ? warps : 32), iters = 2000;
        k<MODE><<<sms, threads>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        k<MODE><<<sms, threads>>>(out, iters, cyc);
        cudaDeviceSynchronize();
        long long h[1024];
        cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i];
        avg /= sms;
        printf("%-28s warps/SM=%2d  cycles/iter=%8.1f  warp-ops/clk/SM=%6.3f  cycles per op per warp=%6.1f\n", name, warps,
               avg / iters, (double)ops_per_iter * warps / (avg / iters), avg / iters / ops_per_iter);
    }
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 1024);
    run<0>("DFMA x8 independent", 8, out, cyc, sms);
    run<1>("DFMA dependent chain", 8, out, cyc, sms);
    run<5>("DADD x8 independent", 8, out, cyc, sms);
    run<6>("DSETP+DADD+SEL x8", 8, out, cyc, sms);
    run<2>("__dsqrt_rn x8", 8, out, cyc, sms);
    run<3>("__ddiv_rn x8", 8, out, cyc, sms);
    run<4>("log x8", 8, out, cyc, sms);
    run<7>("4 DFMA + int mix", 4, out, cyc, sms);
    return 0;
}
