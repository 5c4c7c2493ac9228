// ubench3.cu -- does an FFMA2 (fma.rn.f32x2) occupy the issue port for one cycle or two?  Loop body: NF FFMA2 on
// independent accumulators interleaved with NI integer LOP3s (ALU pipe, independent registers) or NS scalar FFMAs.
#include <cuda_runtime.h>
#include <cstdio>
template <int NF, int NI, int NS>
__global__ void __launch_bounds__(512) k(float *out, int iters, float c0, float c1, unsigned x0, long long *cyc)
{
    float2 a[16]; unsigned u[16]; float f[16];
    for (int i = 0; i < 16; i++) { a[i] = make_float2(threadIdx.x * 1e-6f + i, threadIdx.x * 2e-6f - i); u[i] = threadIdx.x + i; f[i] = i + threadIdx.x; }
    const float2 m = make_float2(c0, c0), d = make_float2(c1, c1);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        constexpr int N = NF > NI ? (NF > NS ? NF : NS) : (NI > NS ? NI : NS);
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (i < NF) a[i % 16] = __ffma2_rn(a[i % 16], m, d);
            if (i < NI) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i % 16]) : "r"(x0), "r"(u[(i + 1) % 16]));
            if (i < NS) f[i % 16] = __fmaf_rn(f[i % 16], c0, c1);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 16; i++) s += a[i].x + a[i].y + (float)u[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF, int NI, int NS>
static void run()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, (size_t)sms * 512 * 4);
    long long *cyc; cudaMallocManaged(&cyc, 8);
    const int iters = 4000;
    printf("FFMA2 %2d  LOP3 %2d  FFMA %2d per trip |", NF, NI, NS);
    for (int w = 1; w <= 4; w++) {
        k<NF, NI, NS><<<sms, 128 * w>>>(out, 10, 0.999f, 1e-3f, 0x5a5a5a5au, cyc);
        k<NF, NI, NS><<<sms, 128 * w>>>(out, iters, 0.999f, 1e-3f, 0x5a5a5a5au, cyc);
        cudaDeviceSynchronize();
        printf("  %dw: %.1f clk/trip/warp-slot", w, (double)*cyc / iters / w);
    }
    printf("\n");
    cudaFree(out);
}
int main()
{
    run<32, 0, 0>();
    run<0, 32, 0>();
    run<0, 0, 32>();
    run<32, 32, 0>();
    run<32, 16, 0>();
    run<32, 0, 32>();
    run<16, 0, 32>();
    return 0;
}
