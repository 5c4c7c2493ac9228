// homography.cu -- plane-induced homographies (prologue) and the un-fused warp operators.
//
// Restates, for the GPU, the host-side maths of HomographySample.sample / sample_inverse
// (utils/mpi/homography_sampler.py:96-122, :176-197) and the per-pixel part (:124-158, :199-220).
#include "homography.cuh"

__global__ void plane_homographies_kernel(const float *__restrict__ G, const float *__restrict__ Kinv,
                                          const float *__restrict__ Kt, const float *__restrict__ depth,
                                          int N, float *__restrict__ Hts, float *__restrict__ Hst,
                                          int *__restrict__ status)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float g[16], ki[9], kt[9], hts[9], hst[9];
    for (int i = 0; i < 16; i++) g[i] = G[(size_t)n * 16 + i];
    for (int i = 0; i < 9; i++) { ki[i] = Kinv[(size_t)n * 9 + i]; kt[i] = Kt[(size_t)n * 9 + i]; }
    const bool ok = plane_homography(g, ki, kt, depth[n], hts, hst);
    for (int i = 0; i < 9; i++) { Hts[(size_t)n * 9 + i] = hts[i]; Hst[(size_t)n * 9 + i] = hst[i]; }
    if (!ok && status) atomicExch(status, 1);
}

__global__ void __launch_bounds__(256)
homography_sample_kernel(const float *__restrict__ src, const float *__restrict__ Hst, int N, int C,
                         int H, int W, float *__restrict__ tgt, uint8_t *__restrict__ valid,
                         float *__restrict__ flowB2A)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (x >= W || y >= H) return;
    float h[9];
#pragma unroll
    for (int i = 0; i < 9; i++) h[i] = __ldg(Hst + (size_t)n * 9 + i);
    const float Wf = (float)W, Hf = (float)H;
    const SrcPos p = tgt_to_src(h, (float)x, (float)y, Wf, Hf, __frcp_rn(Wf * 0.5f), __frcp_rn(Hf * 0.5f));
    const size_t HW = (size_t)H * W, q = (size_t)y * W + x;
    if (valid) valid[(size_t)n * HW + q] = p.valid ? 1 : 0;
    if (flowB2A) {
        float2 f = make_float2(__fsub_rn(p.sx, (float)x), __fsub_rn(p.sy, (float)y));  // :139-141
        reinterpret_cast<float2 *>(flowB2A)[(size_t)n * HW + q] = f;
    }
    const float x0f = floorf(p.ix), y0f = floorf(p.iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);  // weight is 0 where clamped
    const float wx1 = __fsub_rn(__fadd_rn(x0f, 1.f), p.ix), wx0 = __fsub_rn(p.ix, x0f);
    const float wy1 = __fsub_rn(__fadd_rn(y0f, 1.f), p.iy), wy0 = __fsub_rn(p.iy, y0f);
    const float nw = __fmul_rn(wx1, wy1), ne = __fmul_rn(wx0, wy1);
    const float sw = __fmul_rn(wx1, wy0), se = __fmul_rn(wx0, wy0);
    const size_t o00 = (size_t)y0 * W + x0, o01 = (size_t)y0 * W + x1;
    const size_t o10 = (size_t)y1 * W + x0, o11 = (size_t)y1 * W + x1;
    for (int c = 0; c < C; c++) {
        const float *img = src + ((size_t)n * C + c) * HW;
        float acc = __fmul_rn(__ldg(img + o00), nw);
        acc = __fmaf_rn(__ldg(img + o01), ne, acc);
        acc = __fmaf_rn(__ldg(img + o10), sw, acc);
        acc = __fmaf_rn(__ldg(img + o11), se, acc);
        tgt[((size_t)n * C + c) * HW + q] = acc;
    }
}

// a5: HomographySample.sample_inverse (:199-220)
__global__ void __launch_bounds__(256)
homography_flow_kernel(const float *__restrict__ Hts, int N, int H, int W, float *__restrict__ flowA2B)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (x >= W || y >= H) return;
    float h[9];
#pragma unroll
    for (int i = 0; i < 9; i++) h[i] = __ldg(Hts + (size_t)n * 9 + i);
    const float xf = (float)x, yf = (float)y;
    const float X = mf_dot3(h[0], h[1], h[2], xf, yf);
    const float Y = mf_dot3(h[3], h[4], h[5], xf, yf);
    const float Z = mf_dot3(h[6], h[7], h[8], xf, yf);
    float tx, ty;
    mf_div2(X, Y, Z, tx, ty);
    reinterpret_cast<float2 *>(flowA2B)[((size_t)n * H + y) * W + x] =
        make_float2(__fsub_rn(tx, xf), __fsub_rn(ty, yf));
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
MPIFLOW_API int mpiflow_plane_homographies_f32(const float *G, const float *Kinv, const float *Kt,
                                               const float *depth, int N, float *Hts, float *Hst,
                                               int *status, void *stream)
{
    MF_CHECK_ARG(G && Kinv && Kt && depth && Hts && Hst && N > 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (status) {
        cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
        if (e != cudaSuccess) return (int)e;
    }
    plane_homographies_kernel<<<(N + 127) / 128, 128, 0, st>>>(G, Kinv, Kt, depth, N, Hts, Hst, status);
    return mf_launch_status();
}

MPIFLOW_API int mpiflow_homography_sample_f32(const float *src, const float *Hst, int N, int C, int H,
                                              int W, float *tgt, uint8_t *valid, float *flowB2A,
                                              void *stream)
{
    MF_CHECK_ARG(src && Hst && tgt && N > 0 && C > 0 && H > 0 && W > 0 && N <= 65535);
    dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    homography_sample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, Hst, N, C, H, W, tgt, valid, flowB2A);
    return mf_launch_status();
}

MPIFLOW_API int mpiflow_homography_flow_f32(const float *Hts, int N, int H, int W, float *flowA2B,
                                            void *stream)
{
    MF_CHECK_ARG(Hts && flowA2B && N > 0 && H > 0 && W > 0 && N <= 65535);
    dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    homography_flow_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Hts, N, H, W, flowA2B);
    return mf_launch_status();
}
