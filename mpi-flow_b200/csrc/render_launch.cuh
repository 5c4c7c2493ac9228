// render_launch.cuh -- launcher of the persistent renderer, shared by the translation units that instantiate its
// shapes (render_pk16.cu: 64x16 tiles, render_pk8.cu: 64x8 tiles; compiled in parallel).
#pragma once
#include "render_pk.cuh"

struct RenderMaps {
    CUtensorMap rgb, sig, obj, sigd, dist;
};
// tensor maps over the caller's planar tensors for a box bw x bh and a tile tw x th (render_fused.cu)
bool mf_make_render_maps(const RenderParams &p, uint32_t bw, uint32_t bh, uint32_t tw, uint32_t th, RenderMaps *m);
int mf_sm_count();

// one CTA per SM, tiles from the counter the prologue reset; *used = false when the shape does not fit (shared memory,
// tensors TMA cannot describe)
template <class Cfg>
static int launch_pk(const RenderParams &p, int *tile_counter, cudaStream_t st, bool *used)
{
    *used = false;
    const size_t smem = Cfg::smem_bytes(p.S);
    if (smem > 227 * 1024) return 0;
    RenderMaps m;
    if (!mf_make_render_maps(p, Cfg::BW, Cfg::BH, Cfg::TW, Cfg::TH, &m)) return 0;
    const bool hard = (p.flags & MPIFLOW_HARD_FLOW) != 0;
    auto kern = p.obj ? (hard ? render_fused_pk_kernel<Cfg, true, true> : render_fused_pk_kernel<Cfg, true, false>)
                      : (hard ? render_fused_pk_kernel<Cfg, false, true> : render_fused_pk_kernel<Cfg, false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long ntiles = (long long)((p.W + Cfg::TW - 1) / Cfg::TW) * ((p.H + Cfg::TH - 1) / Cfg::TH) * p.B;
    const int sms = mf_sm_count() - p.reserve_sms > 0 ? mf_sm_count() - p.reserve_sms : 1;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    kern<<<grid, Cfg::THREADS, smem, st>>>(m.rgb, m.sig, m.obj, m.sigd, m.dist, p, tile_counter);
    *used = true;
    return mf_launch_status();
}

// 64x16 tiles, ring of `stages` (render_pk16.cu); 64x8 tiles (render_pk8.cu); development shapes (MPIFLOW_AB builds)
int mf_launch_pk16(int stages, const RenderParams &p, int *tile_counter, cudaStream_t st, bool *used);
int mf_launch_pk8(int stages, const RenderParams &p, int *tile_counter, cudaStream_t st, bool *used);
int mf_launch_pk_ab(int which, const RenderParams &p, int *tile_counter, cudaStream_t st, bool *used);
