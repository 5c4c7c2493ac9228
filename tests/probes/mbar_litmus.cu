// mbar_litmus.cu -- the hand-off pattern of render_pk.cuh in isolation: a producer warp fills a shared-memory slot with
// generic stores and publishes it with mbarrier.arrive (release); a consumer warp waits on the barrier (try_wait,
// acquire), reads the slot and hands it back through a second barrier.  compute-sanitizer's racecheck does not model
// mbarrier ordering and reports these stores / loads as hazards; this probe shows what the ordering is worth:
//   mbar_litmus 0   correct protocol                     -> 0 stale reads over 200 k hand-offs
//   mbar_litmus 1   payload written AFTER the arrive     -> stale reads (the test expects > 0)
//   mbar_litmus 2   consumer does not wait               -> stale reads
// usage: mbar_litmus <mode> ; prints "stale <count> of <reads>"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 ::"r"(s32(b)), "r"(parity) : "memory");
}

constexpr int SLOTS = 2, WORDS = 256;

__global__ void litmus(int mode, int iters, unsigned long long *stale, unsigned long long *reads)
{
    __shared__ uint32_t slot[SLOTS][WORDS];
    __shared__ __align__(8) uint64_t full[SLOTS], empty[SLOTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < SLOTS; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < SLOTS * WORDS; i += blockDim.x) (&slot[0][0])[i] = 0xffffffffu;
    __syncthreads();
    unsigned long long bad = 0, n = 0;
    for (int it = 0; it < iters; it++) {
        const int s = it % SLOTS;
        const uint32_t round = it / SLOTS;
        if (warp == 0) {                                   // producer (the planning warp of the renderer)
            if (round) mbar_wait(&empty[s], (round - 1) & 1);
            if (mode == 1) {                               // broken: publish first, write later
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                __nanosleep(200);
            }
            for (int w = lane; w < WORDS; w += 32) slot[s][w] = (uint32_t)it;
            if (mode != 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        } else {                                           // consumer
            if (mode != 2) mbar_wait(&full[s], round & 1);
            for (int w = lane; w < WORDS; w += 32) {
                const uint32_t v = *(volatile uint32_t *)&slot[s][w];
                bad += v != (uint32_t)it;
                n++;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (mode == 2) __nanosleep(20);
        }
    }
    if (warp == 1) { atomicAdd(stale, bad); atomicAdd(reads, n); }
}

int main(int argc, char **argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    unsigned long long *d, h[2] = {0, 0};
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    litmus<<<64, 64>>>(mode, mode == 2 ? 2000 : 200000, d, d + 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("stale %llu of %llu\n", h[0], h[1]);
    return 0;
}
