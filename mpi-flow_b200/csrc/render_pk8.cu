// render_pk8.cu -- the persistent renderer on 64x8 tiles (box 80x12), one pixel pair per thread, 8 + 2 warps: small
// launches (few 64x16 tiles per SM), where the halved tile keeps the persistent CTAs balanced (render_launch.cuh).
#include "render_launch.cuh"

int mf_launch_pk8(int stages, const RenderParams &p, int *tile_counter, cudaStream_t st, bool *used)
{
    if (stages >= 8) return launch_pk<PkCfg<64, 8, 80, 12, 8, 1>>(p, tile_counter, st, used);
    return launch_pk<PkCfg<64, 8, 80, 12, 4, 1>>(p, tile_counter, st, used);
}
