// render_common.cuh -- parameter block, PTX wrappers (mbarrier, TMA, shared-window loads) and per-pixel state shared by
// the fused renderers (render_pk.cuh: persistent TMA kernel; render_fused.cu: generic kernel).  sm_100a only.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "homography.cuh"

struct RenderParams {
    const float *rgb;   int64_t rgb_sb, rgb_ss, rgb_sc;
    const float *sigma; int64_t sig_sb, sig_ss;
    const float *obj;   int64_t obj_sb, obj_ss;
    const PlaneConst *pc;  // [B*S]
    const BatchConst *bc;  // [B]
    int B, S, H, W;
    unsigned flags;
    float *out_rgb, *out_depth, *out_mask, *out_flow, *out_obj;
    const float *spacing_table;  // [S,H,W] source-frame plane spacing of batch entry 0, or nullptr (render_pk.cuh)
    int poses;                   // consecutive frames per MPI (MPIFLOW_POSES_PER_MPI), >= 1
    int reserve_sms;             // SMs the persistent kernel leaves free (MPIFLOW_RESERVE_SMS)
    float hW, hH, rW, rH;        // W / 2, H / 2 and their fp32 reciprocals (grid normalisation, homography_sampler.py:151-154)
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2, int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// shared-memory load at a 32-bit shared-window address + compile-time byte offset (keeps the tap
// addressing to one register and immediates; no generic->shared conversion inside the plane loop)
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ float fast_sqrt(float v)
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float fast_ex2(float v)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

constexpr float MF_LOG2E = 1.4426950408889634f;

// ---- per-pixel state ---------------------------------------------------------------------------------
// front-to-back accumulators of one pixel (mpi_rendering.py:79-96, :143-152)
struct PixelAcc {
    float acc, r, g, b, zsum, wsum, obj, cnt;  // target frame
    float acc_s, fu, fv, fub, fvb, best_w;     // source frame: flow total and the sum of the current block of four planes
};
// the plane that is still waiting for its successor's geometry (general path only): its thickness is
// |xyz[s+1] - xyz[s]| evaluated at the two clamped sample positions (:68-71)
struct Pending {
    float r, g, b, sig, m, x, y, z;
};

