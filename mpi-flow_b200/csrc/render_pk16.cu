// render_pk16.cu -- the persistent renderer on 64x16 tiles (box 80x20: zoom up to ~1.19), one pixel pair per thread,
// 16 + 2 warps: the shape of large launches.  Ring of 5 stages up to S = 64 planes, 4 beyond (render_launch.cuh).
#include "render_launch.cuh"

int mf_launch_pk16(int stages, const RenderParams &p, int *tile_counter, cudaStream_t st, bool *used)
{
#ifdef MPIFLOW_AB
    // stages >= 10: the split shape (16 target-frame + 8 source-frame + 2 producer warps), ring of stages - 10; measured
    // no faster than the default (profiles/r02_experiments.txt), so only development builds instantiate it
    if (stages >= 15) return launch_pk<PkCfg<64, 16, 80, 20, 5, 1, 1>>(p, tile_counter, st, used);
    if (stages >= 10) return launch_pk<PkCfg<64, 16, 80, 20, 4, 1, 1>>(p, tile_counter, st, used);
#endif
    if (stages >= 5) return launch_pk<PkCfg<64, 16, 80, 20, 5, 1>>(p, tile_counter, st, used);
    return launch_pk<PkCfg<64, 16, 80, 20, 4, 1>>(p, tile_counter, st, used);
}

int mf_launch_pk_ab(int which, const RenderParams &p, int *tile_counter, cudaStream_t st, bool *used)
{
#ifdef MPIFLOW_AB
    // A/B shapes of development builds (nvcc -DMPIFLOW_AB), selected with MPIFLOW_PK_CFG
    if (which == 1) return launch_pk<PkCfg<64, 16, 80, 20, 5, 2>>(p, tile_counter, st, used);   // 8+2 warps, two pairs per thread
    if (which == 2) return launch_pk<PkCfg<64, 32, 80, 36, 2, 2>>(p, tile_counter, st, used);   // 16+2 warps, two pairs per thread
#endif
    (void)which; (void)p; (void)tile_counter; (void)st;
    *used = false;
    return 0;
}
