// ubench.cu -- issue/pipe throughput probes for the renderer's instruction mix on sm_100a:
//   mode 0: scalar FFMA          mode 1: FFMA2 (fma.rn.f32x2)       mode 2: FFMA2 + LDS.32 (1:1)
//   mode 3: scalar FFMA + LDS.32 (2:1)   mode 4: FFMA2 + MUFU.EX2 (4:1)   mode 5: LDS.32 only
// prints warp-instructions / clk / SM and fp32 lane-ops / clk / SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template <int MODE>
__global__ void __launch_bounds__(1024) k(float *out, int iters, float c0, float c1)
{
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)i * 1e-9f;
    __syncthreads();
    float2 a[8];
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 1e-6f + i, threadIdx.x * 2e-6f - i);
    const float2 m = make_float2(c0, c0), d = make_float2(c1, c1);
    unsigned addr = threadIdx.x & 1023;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, c0, c1); a[i].y = fmaf(a[i].y, c0, c1); }
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, d); }
            if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, make_float2(sm[(addr + i * 32) & 4095], c1)); }
            if (MODE == 3) { a[i].x = fmaf(a[i].x, c0, sm[(addr + i * 32) & 4095]); a[i].y = fmaf(a[i].y, c0, c1); }
            if (MODE == 4) { a[i] = __ffma2_rn(a[i], m, d); if ((i & 3) == 0) a[i].x = exp2f(a[i].x); }
            if (MODE == 5) { a[i].x += sm[(addr + i * 32) & 4095]; }
        }
        addr = (addr + 33) & 4095;
    }
    float s = 0.f;
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
static void run(const char *name, double winstr_per_iter_thread, double lane_ops_per_iter_thread)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *out; cudaMalloc(&out, (size_t)sms * 1024 * 4);
    const int iters = 20000;
    k<MODE><<<sms, 1024>>>(out, 100, 0.999f, 1e-3f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<MODE><<<sms, 1024>>>(out, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double clk = (double)khz * 1e3 * best * 1e-3;   // cycles at the nominal max clock
    const double warps = 32.0;
    printf("%-28s %.3f ms  warp-instr/clk/SM %.2f   fp32 lane-ops/clk/SM %.1f  (at %d MHz nominal)\n", name, best,
           winstr_per_iter_thread * iters * warps / clk, lane_ops_per_iter_thread * iters * 1024 / clk, khz / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("FFMA scalar", 16, 16);
    run<1>("FFMA2", 8, 16);
    run<2>("FFMA2 + LDS 1:1", 16, 16);
    run<3>("FFMA + LDS 2:1", 24, 16);
    run<4>("FFMA2 + MUFU 4:1", 10, 16);
    run<5>("FADD + LDS 1:1", 16, 8);
    return 0;
}
