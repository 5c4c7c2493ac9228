// ubench2.cu -- does shared-memory load traffic overlap with packed-FP32 (FFMA2) work on one SM sub-partition?
// loop body: NF FFMA2 (independent accumulators) + NL LDS.32 (conflict-free, fixed addresses) [+ NM MUFU.EX2];
// prints cycles per loop trip per SMSP for 1..4 warps per SMSP next to the per-pipe lower bounds.
#include <cuda_runtime.h>
#include <cstdio>
template <int NF, int NL, int NM>
__global__ void __launch_bounds__(512) k(float *out, int iters, float c0, float c1, long long *cyc)
{
    __shared__ float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i * 1e-9f;
    __syncthreads();
    float2 a[16];
    for (int i = 0; i < 16; i++) a[i] = make_float2(threadIdx.x * 1e-6f + i, threadIdx.x * 2e-6f - i);
    const float2 m = make_float2(c0, c0);
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
    float l[16];
    for (int i = 0; i < 16; i++) l[i] = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NL; i++)
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l[i % 16]) : "r"(base + i * 128));
#pragma unroll
        for (int i = 0; i < NF; i++) a[i % 16] = __ffma2_rn(a[i % 16], m, make_float2(l[i % 16], c1));
#pragma unroll
        for (int i = 0; i < NM; i++) a[i % 16].x = exp2f(a[i % 16].x);
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 16; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF, int NL, int NM>
static void run()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, (size_t)sms * 512 * 4);
    long long *cyc; cudaMallocManaged(&cyc, 8);
    const int iters = 4000;
    printf("FFMA2 %2d  LDS %2d  MUFU %2d per trip | bounds/SMSP/warp: fma %d clk, lds(SM-wide x4) %d clk, issue %d |", NF, NL, NM, 2 * NF, NL * 4, NF + NL + NM);
    for (int w = 1; w <= 4; w++) {
        k<NF, NL, NM><<<sms, 128 * w>>>(out, 10, 0.999f, 1e-3f, cyc);
        k<NF, NL, NM><<<sms, 128 * w>>>(out, iters, 0.999f, 1e-3f, cyc);
        cudaDeviceSynchronize();
        printf("  %dw/SMSP: %.1f clk/trip/SMSP", w, (double)*cyc / iters);
    }
    printf("\n");
    cudaFree(out);
}
int main()
{
    run<32, 0, 0>();
    run<0, 16, 0>();
    run<32, 16, 0>();
    run<32, 16, 3>();
    run<48, 24, 4>();
    run<32, 8, 0>();
    run<16, 16, 0>();
    return 0;
}
