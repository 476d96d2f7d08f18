// FP64 FMA peak micro-benchmark (the denominator of the FP64-pipe roofline; MEASURED_PEAKS.json
// has no FP64 figure). Each thread runs ILP independent DFMA chains; reports TFLOP/s and the
// DFMA issue rate per SM per clock for several occupancies.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma(double* out, int iters, double a, double b)
{
    double r[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) r[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += r[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
void run(int ctas_per_sm, int threads, int sms, double* d)
{
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma<ILP><<<sms * ctas_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma<ILP><<<sms * ctas_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double fmas = (double)sms * ctas_per_sm * threads * (double)iters * ILP;
    printf("{\"ilp\": %d, \"ctas_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"dfma_per_s\": %.4e, \"tflops\": %.2f}\n", ILP, ctas_per_sm,
           threads, best, fmas / (best * 1e-3), 2 * fmas / (best * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, clk);
    double* d;
    cudaMalloc(&d, 8);
    run<8>(1, 128, p.multiProcessorCount, d);
    run<8>(2, 256, p.multiProcessorCount, d);
    run<8>(4, 256, p.multiProcessorCount, d);
    run<4>(4, 256, p.multiProcessorCount, d);
    run<2>(8, 256, p.multiProcessorCount, d);
    run<1>(8, 256, p.multiProcessorCount, d);
    return 0;
}
