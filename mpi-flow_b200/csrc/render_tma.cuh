// render_tma.cuh -- TMA-staged variant of the fused MPI renderer (sm_100a).
//
// CTA = one TW x TH tile of pixel indices of one batch entry.  For every plane s one warp (which also
// renders) maps the tile corners through H_src_tgt(s) (same arithmetic as the pixels), derives the
// source footprint, and issues cp.async.bulk.tensor loads (rgb x3, sigma, object mask) of a
// fixed BW x BH box around it into a STAGES-deep shared-memory ring guarded by full/empty mbarriers
// (the producer role rotates over the warps, plane by plane).
// Consumer warps (PPT pixels per thread, same column, rows TH/PPT apart) take their four bilinear
// taps per channel from shared memory with immediate offsets, so the plane loop has no 64-bit address
// arithmetic and HBM/L2 latency is hidden by the ring instead of by occupancy.
// A plane whose footprint does not fit the box (strong zoom / degenerate pose) is read with plain
// global loads for that (tile, plane) only -- results are identical.
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

// ---- per-pixel state ---------------------------------------------------------------------------------
// front-to-back accumulators of one pixel (mpi_rendering.py:79-96, :143-152)
struct PixelAcc {
    float acc, r, g, b, zsum, wsum, obj, cnt;  // target frame
    float acc_s, fu, fv, best_w;               // source frame
};
// the plane that is still waiting for its successor's geometry (general path only): its thickness is
// |xyz[s+1] - xyz[s]| evaluated at the two clamped sample positions (:68-71)
struct Pending {
    float r, g, b, sig, m, x, y, z;
};

template <int TW_, int TH_, int BW_, int BH_, int STAGES_, int PPT_, int MINB_ = 1>
struct TmaCfg {
    static constexpr int MINB = MINB_;                    // CTAs per SM the register budget is sized for
    static constexpr int TW = TW_, TH = TH_, BW = BW_, BH = BH_, STAGES = STAGES_, PPT = PPT_;
    static constexpr int ROWS = TH / PPT;                 // rows per pass
    static constexpr int CONSUMER_WARPS = (TW / 32) * ROWS;
    static constexpr int THREADS = CONSUMER_WARPS * 32;   // warp 0 doubles as the TMA producer
    static constexpr int CH_FLOATS = BW * BH;             // one channel of one stage
    static constexpr int DIRECT_FLOATS = TW * TH;         // un-warped sigma of the tile itself
    static constexpr int STAGE_FLOATS = 5 * CH_FLOATS + DIRECT_FLOATS;  // r,g,b,sigma,obj boxes + direct sigma
    static constexpr size_t STAGE_BYTES = (size_t)STAGE_FLOATS * sizeof(float);
    static_assert(TW % 32 == 0 && TH % PPT == 0 && BW % 4 == 0, "tile shape");
    static_assert((CH_FLOATS * 4) % 128 == 0 && (DIRECT_FLOATS * 4) % 128 == 0,
                  "TMA destinations must stay 128-byte aligned");
    static size_t smem_bytes(int S)
    {
        return STAGES * STAGE_BYTES + (size_t)S * sizeof(PlaneConst) + 2 * STAGES * 8 + 2 * STAGES * 16;
    }
};

constexpr float MF_LOG2E = 1.4426950408889634f;

template <class Cfg, bool HAS_OBJ, bool HARD>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
render_fused_tma_kernel(const __grid_constant__ CUtensorMap map_rgb, const __grid_constant__ CUtensorMap map_sig,
                        const __grid_constant__ CUtensorMap map_obj, const __grid_constant__ CUtensorMap map_sigd,
                        const RenderParams p)
{
    constexpr int TW = Cfg::TW, TH = Cfg::TH, BW = Cfg::BW, BH = Cfg::BH, STAGES = Cfg::STAGES, PPT = Cfg::PPT;
    constexpr int CH = Cfg::CH_FLOATS;
    // no integer round-trip on the pointer: the compiler must keep seeing the shared address space
    // (generic LD instead of LDS otherwise); TMA needs the 128-byte alignment, checked below
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *stage_buf = reinterpret_cast<float *>(smem_raw);
    PlaneConst *spc = reinterpret_cast<PlaneConst *>(smem_raw + STAGES * Cfg::STAGE_BYTES);
    const int S = p.S, H = p.H, W = p.W;
    // the plan of plane sl is written when its stage is free, i.e. after everybody has read the plan of
    // plane sl-STAGES; planes in between are still live: 2*STAGES slots keep them apart
    constexpr int PLAN_SLOTS = 2 * STAGES;
    int4 *meta = reinterpret_cast<int4 *>(spc + S);                              // plans {bx, by, fits, -}
    uint64_t *full = reinterpret_cast<uint64_t *>(meta + PLAN_SLOTS);            // "stage filled" mbarriers
    uint64_t *empty = full + STAGES;                                             // "stage consumed" mbarriers

    const int b = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    {   // plane constants of this batch entry + barrier init
        const float4 *src = reinterpret_cast<const float4 *>(p.pc + (size_t)b * S);
        float4 *dst = reinterpret_cast<float4 *>(spc);
        for (int i = threadIdx.x; i < S * 8; i += Cfg::THREADS) dst[i] = __ldg(src + i);
        if (threadIdx.x == 0) {
            if (smem_u32(smem_raw) & 127u) __trap();
            for (int i = 0; i < STAGES; i++) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], Cfg::CONSUMER_WARPS);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();

    const float Wf = (float)W, Hf = (float)H;
    const float rW = __frcp_rn(Wf * 0.5f), rH = __frcp_rn(Hf * 0.5f);

    // ---- producer role: plan(sl) derives the footprint of plane sl for this tile (lanes 0..3 own the four
    // clamped-to-image tile corners), issue(sl) fires the TMA loads into stage sl % STAGES ----
    const int cx = min(tx0 + ((lane & 1) ? TW - 1 : 0), W - 1);
    const int cy = min(ty0 + ((lane & 2) ? TH - 1 : 0), H - 1);
    auto plan = [&](int sl) {
        const PlaneConst &c = spc[sl];
        const SrcPos sp = tgt_to_src(c.hst, (float)cx, (float)cy, Wf, Hf, rW, rH);
        const float Zc = mf_dot3(c.hst[6], c.hst[7], c.hst[8], (float)cx, (float)cy);
        float mnx = sp.ix, mxx = sp.ix, mny = sp.iy, mxy = sp.iy;
        int ok = (Zc > 0.f) && (sp.ix == sp.ix) && (sp.iy == sp.iy);
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
            ok &= __shfl_xor_sync(0xffffffffu, ok, o);
        }
        if (lane == 0) {
            // 0.02 px of slack covers fp32 rounding of interior pixels relative to the corners;
            // TMA needs a 16-byte aligned box start in the innermost dimension: round x down to 4 texels
            const int bx = ((int)floorf(mnx - 0.02f)) & ~3, by = (int)floorf(mny - 0.02f);
            const int ex = (int)floorf(mxx + 0.02f) + 1, ey = (int)floorf(mxy + 0.02f) + 1;  // last tap col/row
            const bool fits = ok && (ex - bx + 1 <= BW) && (ey - by + 1 <= BH);
            meta[sl % PLAN_SLOTS] = make_int4(bx, by, fits ? 1 : 0, 0);
        }
        __syncwarp();
    };
    auto issue = [&](int sl) {  // one thread
        const int st = sl % STAGES;
        const int4 mt = meta[sl % PLAN_SLOTS];
        float *dst = stage_buf + (size_t)st * Cfg::STAGE_FLOATS;
        const uint32_t box_bytes = mt.z ? (HAS_OBJ ? 5 : 4) * CH * (uint32_t)sizeof(float) : 0u;
        mbar_arrive_expect_tx(&full[st], box_bytes + Cfg::DIRECT_FLOATS * (uint32_t)sizeof(float));
        // the tile's own (un-warped) sigma for the source-frame weights, mpi_rendering.py:104-120
        tma_load_4d(dst + 5 * CH, &map_sigd, &full[st], tx0, ty0, sl, p.sig_sb ? b / p.poses : 0);
        if (mt.z) {
            tma_load_5d(dst, &map_rgb, &full[st], mt.x, mt.y, 0, sl, p.rgb_sb ? b / p.poses : 0);
            tma_load_4d(dst + 3 * CH, &map_sig, &full[st], mt.x, mt.y, sl, p.sig_sb ? b / p.poses : 0);
            if (HAS_OBJ) tma_load_4d(dst + 4 * CH, &map_obj, &full[st], mt.x, mt.y, p.obj_ss ? sl : 0, p.obj_sb ? b : 0);
        }
    };
    for (int sl = 0; sl < STAGES - 1 && sl < S; sl++)
        if (warp == sl % Cfg::CONSUMER_WARPS) {
            plan(sl);
            if (lane == 0) issue(sl);
        }

    // ---- per-thread pixels: same column, rows ROWS apart (out-of-image threads recompute a border pixel) ----
    const int xi = min(tx0 + (warp % (TW / 32)) * 32 + lane, W - 1);
    const int row0 = warp / (TW / 32);
    const float xf = (float)xi;
    const BatchConst &bc = p.bc[b];
    const size_t HW = (size_t)H * W;
    const float *rgb_b = p.rgb + (size_t)b * p.rgb_sb;
    const float *sig_b = p.sigma + (size_t)b * p.sig_sb;
    const float *obj_b = HAS_OBJ ? p.obj + (size_t)b * p.obj_sb : nullptr;

    float yf[PPT], ray0[PPT], ray1[PPT], ray2[PPT], tdist[PPT], invb[PPT];
    uint32_t sd_off[PPT];  // byte offset of the pixel's un-warped sigma inside a stage
    const float bc_a = bc.a;
    const uint32_t stage0_u32 = smem_u32(stage_buf);
    PixelAcc ps[PPT];
    bool slow = false;
#pragma unroll
    for (int j = 0; j < PPT; j++) {
        const int yi = min(ty0 + row0 + j * Cfg::ROWS, H - 1);
        yf[j] = (float)yi;
        sd_off[j] = (uint32_t)(5 * CH + (yi - ty0) * TW + (xi - tx0)) * 4u;
        // source-frame ray |Kinv (x,y,1)|: spacing of planes s, s+1 along it = vnorm * (d[s+1]-d[s])
        const float v0 = mf_dot3(bc.kinv[0], bc.kinv[1], bc.kinv[2], xf, yf[j]);
        const float v1 = mf_dot3(bc.kinv[3], bc.kinv[4], bc.kinv[5], xf, yf[j]);
        const float v2 = mf_dot3(bc.kinv[6], bc.kinv[7], bc.kinv[8], xf, yf[j]);
        ray0[j] = v0; ray1[j] = v1; ray2[j] = v2;
        // target-frame ray r = Kt^-1 (x,y,1); plane s is hit at depth Z_s = (d_s + a) / b(q) with
        // a = (R^T t)_z, b(q) = (R^T r)_z; consecutive hits are |r| |d[s+1]-d[s]| / |b| apart
        const float r0 = mf_dot3(bc.ktinv[0], bc.ktinv[1], bc.ktinv[2], xf, yf[j]);
        const float r1 = mf_dot3(bc.ktinv[3], bc.ktinv[4], bc.ktinv[5], xf, yf[j]);
        const float r2 = mf_dot3(bc.ktinv[6], bc.ktinv[7], bc.ktinv[8], xf, yf[j]);
        const float bq = fmaf(bc.rt2[2], r2, fmaf(bc.rt2[1], r1, bc.rt2[0] * r0));
        invb[j] = __frcp_rn(bq);
        tdist[j] = sqrtf(fmaf(r2, r2, fmaf(r1, r1, r0 * r0))) * fabsf(invb[j]);
        // The sample positions of all planes lie on one (epipolar) segment between those of the first
        // and the last plane, so if neither end is clamped by the border no plane is, and the plane
        // geometry along the pixel's own ray (above) is what the reference's warped xyz amounts to.
        const SrcPos e0 = tgt_to_src(spc[0].hst, xf, yf[j], Wf, Hf, rW, rH);
        const SrcPos e1 = tgt_to_src(spc[S - 1].hst, xf, yf[j], Wf, Hf, rW, rH);
        const bool inside = e0.sx >= 0.f && e0.sx <= Wf - 1.f && e0.sy >= 0.f && e0.sy <= Hf - 1.f &&
                            e1.sx >= 0.f && e1.sx <= Wf - 1.f && e1.sy >= 0.f && e1.sy <= Hf - 1.f;
        slow = slow || !inside || !(bq > 1e-6f) || !(tdist[j] < 1e30f);
        PixelAcc &q = ps[j];
        q.acc = 1.f; q.r = q.g = q.b = q.zsum = q.wsum = q.obj = q.cnt = 0.f;
        q.acc_s = 1.f; q.fu = q.fv = 0.f; q.best_w = -1.f;
    }
    slow = __any_sync(0xffffffffu, slow);  // warp-uniform choice of the plane loop

    // `GENERAL` keeps one plane pending and measures thickness between clamped sample positions
    // (border warps); otherwise thickness and depth come from the pixel's ray and nothing is pending.
    auto plane_loop = [&](auto general_tag) {
        constexpr bool GENERAL = decltype(general_tag)::value;
        Pending pd[GENERAL ? PPT : 1];
        if (GENERAL) {
#pragma unroll
            for (int j = 0; j < PPT; j++) pd[j] = Pending{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        }
        auto composite = [&](PixelAcc &q, float sg, float dist_l2e, float cr, float cg, float cb, float z, float m) {
            const float T = fast_ex2(-sg * dist_l2e);  // exp(-sigma*dist) (:79)
            const float w = q.acc * (1.f - T);         // :80, :89
            q.acc *= (T + 1e-6f);                      // :84
            q.wsum += w;
            q.r = fmaf(w, cr, q.r);
            q.g = fmaf(w, cg, q.g);
            q.b = fmaf(w, cb, q.b);
            q.zsum = fmaf(w, z, q.zsum);
            if (HAS_OBJ) q.obj = fmaf(w, m, q.obj);
        };
        for (int s = 0; s < S; s++) {
            const int st = s % STAGES;
            // The producer role rotates over the warps (plane sl is fetched by warp sl % #warps): a fixed
            // producer warp would be permanently behind by its extra work, and since nobody can run more
            // than STAGES-1 planes ahead of the producer the whole CTA would run at that warp's
            // single-warp speed.
            if (s + STAGES - 1 < S && warp == (s + STAGES - 1) % Cfg::CONSUMER_WARPS) {
                // the stage plane s-1 used is free once every warp has arrived on its empty barrier
                if (s >= 1) mbar_wait(&empty[(s - 1) % STAGES], ((s - 1) / STAGES) & 1);
                plan(s + STAGES - 1);
                if (lane == 0) issue(s + STAGES - 1);
            }
            const PlaneConst &c = spc[s];
            const bool last = (s + 1 == S);
            mbar_wait(&full[st], (s / STAGES) & 1);
            const int4 mt = meta[s % PLAN_SLOTS];
            // 32-bit shared-window addresses: stage base, and the tile base with the box origin folded in
            // (tap address = tile_u32 + 4*(y0*BW + x0) + immediates)
            const uint32_t stage_u32 = stage0_u32 + (uint32_t)st * (uint32_t)Cfg::STAGE_BYTES;
            const uint32_t tile_u32 = stage_u32 - (uint32_t)(mt.y * BW + mt.x) * 4u;
            const float zc = c.depth + bc_a;  // numerator of the plane's depth along the pixel ray
            // thickness of plane s along a ray of unit parameter: d[s+1]-d[s], or 1e3 for the last plane
            // (mpi_rendering.py:73-78); both carry log2(e) for ex2
            const float dmul = last ? 0.f : c.ddepth_l2e, dadd = last ? 1e3f * MF_LOG2E : 0.f;
            // shared by this thread's pixels (same column): first product of each FMA chain
            const float hx0 = __fmul_rn(c.hst[0], xf), hx3 = __fmul_rn(c.hst[3], xf), hx6 = __fmul_rn(c.hst[6], xf);
            const float gx0 = __fmul_rn(c.hts[0], xf), gx3 = __fmul_rn(c.hts[3], xf), gx6 = __fmul_rn(c.hts[6], xf);
            // One straight-line block for all of this thread's pixels (the only branch, "is the footprint
            // in shared memory", is taken outside), so the scheduler can interleave their dependency chains.
            auto body = [&](auto fits_tag) {
                constexpr bool FITS = decltype(fits_tag)::value;
                const float *rgb_s = nullptr, *sig_s = nullptr, *obj_s = nullptr;
                if (!FITS) {
                    int so = s;
                    asm volatile("" : "+r"(so));  // keep these out of the loop's induction variables
                    rgb_s = rgb_b + (size_t)so * p.rgb_ss;
                    sig_s = sig_b + (size_t)so * p.sig_ss;
                    obj_s = HAS_OBJ ? obj_b + (size_t)so * p.obj_ss : nullptr;
                }
#pragma unroll
                for (int j = 0; j < PPT; j++) {
                    PixelAcc &q = ps[j];
                    // ------------- target frame: sample position (homography_sampler.py:131-154) -------------
                    const float X = __fadd_rn(__fmaf_rn(c.hst[1], yf[j], hx0), c.hst[2]);
                    const float Y = __fadd_rn(__fmaf_rn(c.hst[4], yf[j], hx3), c.hst[5]);
                    const float Z = __fadd_rn(__fmaf_rn(c.hst[7], yf[j], hx6), c.hst[8]);
                    float sx, sy;
                    mf_div2(X, Y, Z, sx, sy);
                    if ((sx > -1.f) && (sx < Wf) && (sy > -1.f) && (sy < Hf)) q.cnt += 1.f;  // :143-147, mpi_rendering.py:347
                    const float gx = __fadd_rn(__fmul_rn(__fadd_rn(sx, 0.5f), rW), -1.f);
                    const float gy = __fadd_rn(__fmul_rn(__fadd_rn(sy, 0.5f), rH), -1.f);
                    const float ixu = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), Wf, -1.f), 0.5f);
                    const float iyu = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), Hf, -1.f), 0.5f);
                    const float ix = fminf(Wf - 1.f, fmaxf(ixu, 0.f));
                    const float iy = fminf(Hf - 1.f, fmaxf(iyu, 0.f));
                    // bilinear weights; (x0+1)-ix and 1-(ix-x0) are the same real number, ix-x0 is exact
                    const float x0f = floorf(ix), y0f = floorf(iy);
                    const float wx0 = __fsub_rn(ix, x0f), wy0 = __fsub_rn(iy, y0f);
                    const float wx1 = __fsub_rn(1.f, wx0), wy1 = __fsub_rn(1.f, wy0);
                    const float nw = __fmul_rn(wx1, wy1), ne = __fmul_rn(wx0, wy1);
                    const float sw = __fmul_rn(wx1, wy0), se = __fmul_rn(wx0, wy0);
                    const int x0 = (int)x0f, y0 = (int)y0f;
                    float cr, cg, cb, sg, m = 0.f;
                    if (FITS) {  // the plane's footprint is in shared memory
                        const uint32_t t = tile_u32 + (uint32_t)(y0 * BW + x0) * 4u;
#define MF_TAP(C)                                                                   \
    __fmaf_rn(lds_f32<((C) * CH + BW + 1) * 4>(t), se,                               \
              __fmaf_rn(lds_f32<((C) * CH + BW) * 4>(t), sw,                         \
                        __fmaf_rn(lds_f32<((C) * CH + 1) * 4>(t), ne, __fmul_rn(lds_f32<(C) * CH * 4>(t), nw))))
                        cr = MF_TAP(0); cg = MF_TAP(1); cb = MF_TAP(2); sg = MF_TAP(3);
                        if (HAS_OBJ) m = MF_TAP(4);
#undef MF_TAP
                    } else {     // same taps straight from global memory
                        const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
                        const int o00 = y0 * W + x0, o01 = y0 * W + x1, o10 = y1 * W + x0, o11 = y1 * W + x1;
                        auto tap = [&](const float *img) {
                            float a = __fmul_rn(__ldg(img + o00), nw);
                            a = __fmaf_rn(__ldg(img + o01), ne, a);
                            a = __fmaf_rn(__ldg(img + o10), sw, a);
                            return __fmaf_rn(__ldg(img + o11), se, a);
                        };
                        cr = tap(rgb_s); cg = tap(rgb_s + p.rgb_sc); cb = tap(rgb_s + 2 * p.rgb_sc);
                        sg = tap(sig_s);
                        if (HAS_OBJ) m = tap(obj_s);
                    }
                    // ------------- target frame: plane geometry + compositing -------------
                    if (!GENERAL) {
                        const float Zt = zc * invb[j];
                        sg = (Zt >= 0.f) ? sg : 0.f;  // mpi_rendering.py:334-338
                        composite(q, sg, fmaf(tdist[j], dmul, dadd), cr, cg, cb, Zt, m);
                    } else {
                        // xyz_tgt of the plane at the (possibly clamped) sample position
                        const float Xc = fmaf(c.P[0], ix, fmaf(c.P[1], iy, c.P[2]));
                        const float Yc = fmaf(c.P[3], ix, fmaf(c.P[4], iy, c.P[5]));
                        const float Zt = fmaf(c.P[6], ix, fmaf(c.P[7], iy, c.P[8]));
                        Pending &d = pd[j];
                        if (s > 0) {
                            const float dx = Xc - d.x, dy = Yc - d.y, dz = Zt - d.z;
                            composite(q, d.sig, fast_sqrt(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) * MF_LOG2E, d.r, d.g, d.b, d.z, d.m);
                        }
                        d = Pending{cr, cg, cb, (Zt >= 0.f) ? sg : 0.f, m, Xc, Yc, Zt};
                        if (last) composite(q, d.sig, 1e3f * MF_LOG2E, d.r, d.g, d.b, d.z, d.m);  // :73-78
                    }
                    // ------------- source frame (homography_sampler.py:207-218, mpi_rendering.py:102-133) ------
                    const float U = __fadd_rn(__fmaf_rn(c.hts[1], yf[j], gx0), c.hts[2]);
                    const float V = __fadd_rn(__fmaf_rn(c.hts[4], yf[j], gx3), c.hts[5]);
                    const float Q = __fadd_rn(__fmaf_rn(c.hts[7], yf[j], gx6), c.hts[8]);
                    float tu, tv;
                    mf_div2(U, V, Q, tu, tv);
                    const float du = __fsub_rn(tu, xf), dv = __fsub_rn(tv, yf[j]);
                    const float sgs = lds_f32<0>(stage_u32 + sd_off[j]);  // un-warped sigma of plane s
                    // source-frame thickness and transmittance literally (mpi_rendering.py:104-120): see render_pk.cuh
                    const float ex = __fsub_rn(__fmul_rn(ray0[j], c.depth_next), __fmul_rn(ray0[j], c.depth));
                    const float ey = __fsub_rn(__fmul_rn(ray1[j], c.depth_next), __fmul_rn(ray1[j], c.depth));
                    const float ez = __fsub_rn(__fmul_rn(ray2[j], c.depth_next), __fmul_rn(ray2[j], c.depth));
                    const float ds = last ? 1e3f : sqrtf(__fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex))));
                    const float Ts = expf(__fmul_rn(-sgs, ds));
                    const float ws = q.acc_s * (1.f - Ts);
                    q.acc_s *= (Ts + 1e-6f);
                    if (HARD) {  // one-hot of the first arg-max weight (:126-130)
                        if (ws > q.best_w) { q.best_w = ws; q.fu = du; q.fv = dv; }
                    } else {
                        q.fu = fmaf(ws, du, q.fu);
                        q.fv = fmaf(ws, dv, q.fv);
                    }
                }
            };
            if (mt.z) body(std::true_type{});
            else body(std::false_type{});
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
    };
    if (slow) plane_loop(std::true_type{});
    else plane_loop(std::false_type{});

#pragma unroll
    for (int j = 0; j < PPT; j++) {
        const PixelAcc &q = ps[j];
        const int x = tx0 + (warp % (TW / 32)) * 32 + lane, y = ty0 + row0 + j * Cfg::ROWS;
        if (x >= W || y >= H) continue;
        const size_t o = (size_t)y * W + x;
        float *orgb = p.out_rgb + (size_t)b * 3 * HW + o;
        orgb[0] = q.r; orgb[HW] = q.g; orgb[2 * HW] = q.b;
        p.out_depth[(size_t)b * HW + o] = (p.flags & MPIFLOW_BG_DEPTH_INF) ? fmaf(1.f - q.wsum, 1000.f, q.zsum)
                                                                           : __fdiv_rn(q.zsum, q.wsum + 1e-5f);
        p.out_mask[(size_t)b * HW + o] = q.cnt;
        p.out_flow[(size_t)b * 2 * HW + o] = q.fu;
        p.out_flow[(size_t)b * 2 * HW + HW + o] = q.fv;
        if (HAS_OBJ && p.out_obj) p.out_obj[(size_t)b * HW + o] = q.obj;
    }
}
