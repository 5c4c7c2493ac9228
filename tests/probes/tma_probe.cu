// tma_probe.cu -- which tensor-map shapes does cp.async.bulk.tensor accept on this part?
// usage: tma_probe <rank> <boxw> <boxh> <boxc> <W> <H> <cx> <cy>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, int rank, int cx, int cy, int n, float *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float *buf = (float *)sm;
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(n * 4) : "memory");
        if (rank == 5)
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6,%7}], [%2];"
                         ::"r"(s32(buf)), "l"(&map), "r"(s32(&bar)), "r"(cx), "r"(cy), "r"(0), "r"(1), "r"(1) : "memory");
        else if (rank == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5,%6}], [%2];"
                         ::"r"(s32(buf)), "l"(&map), "r"(s32(&bar)), "r"(cx), "r"(cy), "r"(0), "r"(1) : "memory");
        else if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5}], [%2];"
                         ::"r"(s32(buf)), "l"(&map), "r"(s32(&bar)), "r"(cx), "r"(cy), "r"(0) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4}], [%2];"
                         ::"r"(s32(buf)), "l"(&map), "r"(s32(&bar)), "r"(cx), "r"(cy) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char **argv)
{
    int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), bc = atoi(argv[4]), W = atoi(argv[5]), H = atoi(argv[6]);
    int cx = atoi(argv[7]), cy = atoi(argv[8]);
    const int C = 4, S = 3, B = 2;
    size_t n = (size_t)B * S * C * H * W;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; i++) h[i] = (float)(i % 100003);
    float *d, *out;
    cudaMalloc(&d, n * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    cuuint64_t gd[5] = {(cuuint64_t)W, (cuuint64_t)H, 3, S, B};
    cuuint64_t gs[4] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4, (cuuint64_t)S * C * H * W * 4};
    cuuint32_t box[5] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    if (rank < 3) box[2] = 1;
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    if (r) return 1;
    int cnt = bw * bh * (rank >= 3 ? bc : 1);
    cudaMalloc(&out, cnt * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, cnt * 4 + 128);
    k<<<1, 128, cnt * 4>>>(m, rank, cx, cy, cnt, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e) return 2;
    std::vector<float> o(cnt);
    cudaMemcpy(o.data(), out, cnt * 4, cudaMemcpyDeviceToHost);
    // check element (x=3,y=2,c=0) of the box
    int bad = 0;
    for (int c = 0; c < (rank >= 3 ? bc : 1); c++)
        for (int y = 0; y < bh; y++)
            for (int x = 0; x < bw; x++) {
                int gx = cx + x, gy = cy + y;
                size_t b = rank >= 5 ? 1 : 0, s = rank >= 4 ? 1 : 0;
                float want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[(((b * S + s) * C + c) * H + gy) * W + gx];
                if (o[(c * bh + y) * bw + x] != want) bad++;
            }
    printf("mismatches: %d of %d\n", bad, cnt);
    return 0;
}
