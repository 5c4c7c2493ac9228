// tma_bw.cu -- throughput of cp.async.bulk.tensor box loads (no compute): which box shapes / alignments
// reach HBM speed?  usage: tma_bw <bw> <bh> <xalign> <stages> [ctas_per_sm]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(s32(bar)), "r"(parity) : "memory");
}
// grid: tiles of 64x16 over (W,H); each CTA loops planes s=0..S-1 loading box (bw,bh,4ch as 4D map dims W,H,C*S)
__global__ void k(const __grid_constant__ CUtensorMap map, int S, int bw, int bh, int xalign, int stages, float *sink)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long full[8];
    const unsigned stage_bytes = bw * bh * 4 * 4;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    int bx = ((blockIdx.x * 64 - 3 - (int)(blockIdx.y % 7)) / xalign) * xalign, by = blockIdx.y * 16 - 2;
    float acc = 0.f;
    for (int s = 0; s < S + stages - 1; s++) {
        if (threadIdx.x == 0 && s < S) {
            int st = s % stages;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(stage_bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3,%4,%5}], [%2];"
                         ::"r"(s32(sm + st * stage_bytes)), "l"(&map), "r"(s32(&full[st])), "r"(bx + (s % 5)*xalign), "r"(by), "r"(s * 4) : "memory");
        }
        int c = s - (stages - 1);
        if (c >= 0) {
            wait(&full[c % stages], (c / stages) & 1);
            acc += ((float *)(sm + (c % stages) * stage_bytes))[threadIdx.x];
            __syncthreads();
        }
    }
    if (acc == 1234.5f) sink[0] = acc;
}
int main(int argc, char **argv)
{
    int bw = atoi(argv[1]), bh = atoi(argv[2]), xalign = atoi(argv[3]), stages = atoi(argv[4]);
    const int W = 1280, H = 384, S = 64, C = 4, B = 4;
    size_t n = (size_t)B * S * C * H * W;
    float *d, *sink;
    cudaMalloc(&d, n * 4); cudaMemset(d, 0, n * 4); cudaMalloc(&sink, 4);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    float ms_best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int b = 0; b < B; b++) {
            cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C * S};
            cuuint64_t gs[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
            cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 4}, es[3] = {1, 1, 1};
            CUtensorMap m;
            CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d + (size_t)b * S * C * H * W, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r) { printf("encode rc=%d\n", (int)r); return 1; }
            size_t smem = (size_t)stages * bw * bh * 16;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<dim3(W / 64, H / 16), 128, smem>>>(m, S, bw, bh, xalign, stages, sink);
        }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("kernel: %s\n", cudaGetErrorString(e)); return 2; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < ms_best) ms_best = ms;
    }
    double alg = (double)n * 4 / 1e9, moved = (double)B * (W / 64) * (H / 16) * S * bw * bh * 16 / 1e9;
    printf("box %dx%d xalign %d stages %d: %.3f ms  algorithmic %.0f GB/s  smem-fill %.0f GB/s (overfetch %.2fx)\n", bw, bh, xalign,
           stages, ms_best, alg / (ms_best * 1e-3), moved / (ms_best * 1e-3), moved / alg);
    return 0;
}
