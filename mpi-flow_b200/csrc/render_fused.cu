// render_fused.cu -- the fused MPI renderer (SURVEY.md section 8, rows a2+a3+a4+a5+a7+a8+a10).
//
// One thread per pixel index p = (x, y).  The same index is used twice, as in the reference:
//   * as TARGET pixel q: for every plane s the plane homography H_src_tgt maps q to a source
//     position, the RGBA (+ object mask) plane is sampled there bilinearly with border clamp
//     (homography_sampler.py:124-158 -> F.grid_sample), the plane's target-frame xyz is evaluated at
//     the clamped position, and the stack is over-composited front to back
//     (mpi_rendering.py:62-99, :142-154, :329-347);
//   * as SOURCE pixel p: H_tgt_src gives the per-plane forward flow (homography_sampler.py:199-220),
//     volume-rendered with the weights of the UN-warped sigma (mpi_rendering.py:102-139).
// Running transmittances, colour, depth, flow, mask and count accumulators live in registers for
// the whole plane loop; nothing of size [B,S,*,H,W] is ever written.
//
// Algorithmic HBM bytes per frame: 4*H*W*(4S + 1 + 8)   (RGBA once, mask once, 8 output channels).
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

#include "render_launch.cuh"

// storage types of the MPI in the generic kernel: fp32, or fp16 converted on load (fp32 arithmetic either way)
__device__ __forceinline__ float mf_load(const float *p) { return __ldg(p); }
__device__ __forceinline__ float mf_load(const __half *p) { return __half2float(__ldg(p)); }

// prologue: PlaneConst for every (b, s); inputs are per batch.
__global__ void plane_consts_kernel(const float *__restrict__ G, const float *__restrict__ Kinv,
                                    const float *__restrict__ Kt, const float *__restrict__ disparity,
                                    int B, int S, PlaneConst *__restrict__ pc, BatchConst *__restrict__ bcs,
                                    int *__restrict__ status, int *__restrict__ tile_counter)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0 && tile_counter) *tile_counter = 0;  // work queue of the persistent renderer that follows
    if (n >= B * S) return;
    const int b = n / S, s = n % S;
    float g[16], ki[9], kt[9];
    for (int i = 0; i < 16; i++) g[i] = G[(size_t)b * 16 + i];
    for (int i = 0; i < 9; i++) { ki[i] = Kinv[(size_t)b * 9 + i]; kt[i] = Kt[(size_t)b * 9 + i]; }
    const float depth = __frcp_rn(disparity[n]);  // torch.reciprocal (mpi_rendering.py:278)
    PlaneConst c;
    float hts[9];
    const bool ok = plane_homography(g, ki, kt, depth, hts, c.hst);
    for (int i = 0; i < 3; i++) { c.g2[i] = hts[i * 3 + 2]; c.g01[2 * i] = hts[i * 3]; c.g01[2 * i + 1] = hts[i * 3 + 1]; }
    // xyz_tgt(ix,iy) = G . [ (Kinv (ix,iy,1)) * depth ; 1 ]  ==  (depth * R Kinv) (ix,iy,1) + t
    // (mpi_rendering.py:213-256, rendering_utils.py:4-23; affine in the pixel position, so the
    //  reference's bilinear sample of the xyz tensor equals this evaluated at the clamped index)
    float R[9], M[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i * 3 + j] = g[i * 4 + j];
    mm33_fma(R, ki, M);
    for (int i = 0; i < 9; i++) c.P[i] = __fmul_rn(M[i], depth);
    for (int i = 0; i < 3; i++) c.P[i * 3 + 2] = __fadd_rn(c.P[i * 3 + 2], g[i * 4 + 3]);
    const bool last = s + 1 == S;
    const float depth_next = last ? depth : __frcp_rn(disparity[n + 1]);
    // (R^T t)_z: the target ray of pixel q meets plane s at Z = (depth_s + a) / (R^T Kt^-1 q)_z
    const float a = fmaf(g[2 * 4 + 2], g[2 * 4 + 3], fmaf(g[1 * 4 + 2], g[1 * 4 + 3], g[0 * 4 + 2] * g[0 * 4 + 3]));
    c.depth = depth;
    c.depth_next = depth_next;
    c.zc = depth + a;
    // thickness of the plane along a ray of unit parameter, times log2(e): exp(-sigma*dist) is evaluated as
    // ex2(-sigma*dist*log2e) in the target frame; the last plane is 1e3 thick (mpi_rendering.py:73-78)
    c.nmul = last ? 0.f : __fsub_rn(depth_next, depth) * 1.4426950408889634f;
    c.nadd = last ? 1e3f * 1.4426950408889634f : 0.f;
    pc[n] = c;
    bool okb = true;
    if (s == 0) {
        BatchConst bc;
        double k64[9], ki64[9];
        for (int i = 0; i < 9; i++) { bc.kinv[i] = ki[i]; k64[i] = (double)kt[i]; }
        okb = inverse33_f64(k64, ki64);
        for (int i = 0; i < 9; i++) bc.ktinv[i] = okb ? (float)ki64[i] : 0.f;
        for (int i = 0; i < 3; i++) bc.rt2[i] = g[i * 4 + 2];
        bc.a = a;
        float mind = depth;
        for (int i = 1; i < S; i++) mind = fminf(mind, __frcp_rn(disparity[n + i]));
        bc.min_zc = mind + bc.a;   // NaN-propagating on purpose: a NaN disparity disables the lean loop
        bc.hts_cols_invariant = 0.f;   // set by batch_flags_kernel
        bc.spacing_table = 0.f;        // set by batch_flags_kernel
        bc.pad[0] = bc.pad[1] = bc.pad[2] = 0.f;
        bcs[b] = bc;
    }
    if (!(ok && okb) && status) atomicExch(status, 1);
}

// H_tgt_src = K_t (R + t n^T / d) K_s^-1 with n = (0,0,1): the plane enters through column 2 of the middle factor
// only, so when row 2 of K_s^-1 is (0, 0, 1) columns 0 and 1 of the product do not depend on the plane -- in fp32 as
// well.  Checked on the computed matrices, bit for bit, so that the lean plane loop may hoist h0 x + h1 y.
// Block B (only launched with a table): the key of the plane-spacing table kept in the workspace -- magic, S, H, W,
// K_src^-1 of entry 0 and the plane depths, bit for bit.  key[TABLE_KEY_VALID] = 1 tells spacing_table_kernel that the
// table in the workspace is the one this launch would compute (MPIFLOW_KEEP_TABLE); otherwise the key is rewritten.
constexpr uint32_t TABLE_KEY_MAGIC = 0x4D504954u;
constexpr int TABLE_KEY_VALID = 13, TABLE_KEY_DEPTHS = 16;
__global__ void batch_flags_kernel(const PlaneConst *__restrict__ pc, BatchConst *__restrict__ bcs, int B, int S, int H, int W,
                                   int have_table, int keep, uint32_t *__restrict__ key)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b == B) {
        auto want = [&](int i) -> uint32_t {
            if (i == 0) return TABLE_KEY_MAGIC;
            if (i == 1) return (uint32_t)S;
            if (i == 2) return (uint32_t)H;
            if (i == 3) return (uint32_t)W;
            if (i < 13) return __float_as_uint(bcs[0].kinv[i - 4]);
            if (i < TABLE_KEY_DEPTHS) return 0u;
            return __float_as_uint(pc[i - TABLE_KEY_DEPTHS].depth);
        };
        bool match = keep != 0;
        for (int i = lane; i < TABLE_KEY_DEPTHS + S && match; i += 32)
            if (i != TABLE_KEY_VALID) match = key[i] == want(i);
        match = __all_sync(0xffffffffu, match);
        if (!match)
            for (int i = lane; i < TABLE_KEY_DEPTHS + S; i += 32)
                if (i != TABLE_KEY_VALID) key[i] = want(i);
        if (lane == 0) key[TABLE_KEY_VALID] = match ? 1u : 0u;
        return;
    }
    const PlaneConst *p0 = pc + (size_t)b * S;
    bool same = true, shared = have_table != 0;
    for (int s = lane; s < S; s += 32) {
        for (int i = 0; i < 6; i++) same = same && (__float_as_uint(p0[s].g01[i]) == __float_as_uint(p0[0].g01[i]));
        // source-frame plane spacing depends on K_src^-1 and the plane depths only: the table of entry 0 serves entry b
        shared = shared && (__float_as_uint(p0[s].depth) == __float_as_uint(pc[s].depth)) &&
                 (__float_as_uint(p0[s].depth_next) == __float_as_uint(pc[s].depth_next));
    }
    if (lane < 9) shared = shared && (__float_as_uint(bcs[b].kinv[lane]) == __float_as_uint(bcs[0].kinv[lane]));
    same = __all_sync(0xffffffffu, same);
    shared = __all_sync(0xffffffffu, shared);
    if (lane == 0) {
        bcs[b].hts_cols_invariant = same ? 1.f : 0.f;
        bcs[b].spacing_table = shared ? 1.f : 0.f;
    }
}

constexpr int TABLE_PLANES = 8;   // planes per thread of spacing_table_kernel

// Plane spacing of the source frame for batch entry 0, table[s][y][x] = | fl(ray * depth[s+1]) - fl(ray * depth[s]) |
// (1e3 for the last plane, mpi_rendering.py:104-120), with exactly the operations of the plane loop of render_pk.cuh.
// It depends on K_src^-1 and the plane depths only -- in the generator both are the same for every frame -- so one
// launch computes it once (0.13 GB at 1280x384x64, a few microseconds per frame) instead of per frame and plane.
__global__ void __launch_bounds__(128)
spacing_table_kernel(const PlaneConst *__restrict__ pc, const BatchConst *__restrict__ bcs, int S, int H,
                     int W, float *__restrict__ table, const uint32_t *__restrict__ key)
{
    if (key[TABLE_KEY_VALID]) return;      // the workspace already holds this very table (MPIFLOW_KEEP_TABLE)
    // grid (W / 512, H, S / TABLE_PLANES), W % 4 == 0: four pixels (one 16-byte store) and eight planes per thread --
    // enough stores in flight per thread and few enough blocks (a thread per pixel and four planes was bound by the
    // block launch rate: 61 k blocks, 63 us for 126 MB)
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= W) return;
    const float yf = (float)y;
    const float *k = bcs[0].kinv;
    float r0[4], r1[4], r2[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float xf = (float)(x + i);
        r0[i] = mf_dot3(__ldg(k + 0), __ldg(k + 1), __ldg(k + 2), xf, yf);
        r1[i] = mf_dot3(__ldg(k + 3), __ldg(k + 4), __ldg(k + 5), xf, yf);
        r2[i] = mf_dot3(__ldg(k + 6), __ldg(k + 7), __ldg(k + 8), xf, yf);
    }
    const size_t HW = (size_t)H * W, q = (size_t)y * W + x;
    const int s0 = blockIdx.z * TABLE_PLANES;
#pragma unroll
    for (int u = 0; u < TABLE_PLANES; u++) {
        const int s = s0 + u;
        if (s >= S) break;
        const float dn = __ldg(&pc[s].depth_next), da = __ldg(&pc[s].depth);
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float ex = __fsub_rn(__fmul_rn(r0[i], dn), __fmul_rn(r0[i], da));
            const float ey = __fsub_rn(__fmul_rn(r1[i], dn), __fmul_rn(r1[i], da));
            const float ez = __fsub_rn(__fmul_rn(r2[i], dn), __fmul_rn(r2[i], da));
            // torch.norm over the 3-vector on CUDA: ((x^2 + y^2) + z^2) with every square rounded, then sqrt.rn
            const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
            v[i] = (s + 1 == S) ? 1e3f : sqrtf(fmaxf(e2, 1e-30f));
        }
        *reinterpret_cast<float4 *>(table + (size_t)s * HW + q) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// front-to-back accumulator of one frame (mpi_rendering.py:79-96, :143-152)
struct TgtAcc {
    float acc = 1.f, r = 0.f, g = 0.f, b = 0.f, zsum = 0.f, wsum = 0.f, obj = 0.f;
    __device__ __forceinline__ void add(float sigma, float dist, float cr, float cg, float cb, float z,
                                        float m)
    {
        const float T = __expf(-sigma * dist);  // :79
        const float w = acc * (1.f - T);        // :80, :89
        acc *= (T + 1e-6f);                     // :84
        wsum += w;
        r = fmaf(w, cr, r);
        g = fmaf(w, cg, g);
        b = fmaf(w, cb, b);
        zsum = fmaf(w, z, zsum);
        obj = fmaf(w, m, obj);
    }
};

// TIn = float: tensors TMA cannot describe.  TIn = __half: the fp16-storage entry point (mpiflow_render_fused_f16):
// rgb / sigma / obj of RenderParams then point at IEEE halves (strides in elements).
template <bool HAS_OBJ, class TIn>
__global__ void __launch_bounds__(256)
render_fused_kernel(const RenderParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    PlaneConst *spc = reinterpret_cast<PlaneConst *>(smem_raw);
    const int b = blockIdx.z;
    const int S = p.S, H = p.H, W = p.W;
    {   // stage this batch entry's plane constants: S * 128 B
        const float4 *src = reinterpret_cast<const float4 *>(p.pc + (size_t)b * S);
        float4 *dst = reinterpret_cast<float4 *>(spc);
        for (int i = threadIdx.x; i < S * 8; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;

    const float xf = (float)x, yf = (float)y, Wf = (float)W, Hf = (float)H;
    const float rW = __frcp_rn(Wf * 0.5f), rH = __frcp_rn(Hf * 0.5f);
    const size_t HW = (size_t)H * W, q = (size_t)y * W + x;

    // source-frame ray Kinv (x,y,1): xyz_src = fl(ray * depth) (mpi_rendering.py:234-236); the spacing of
    // consecutive planes is evaluated from those products, like the reference (:104-107)
    float ray0, ray1, ray2;
    {
        const float *k = p.bc[b].kinv;
        ray0 = mf_dot3(__ldg(k + 0), __ldg(k + 1), __ldg(k + 2), xf, yf);
        ray1 = mf_dot3(__ldg(k + 3), __ldg(k + 4), __ldg(k + 5), xf, yf);
        ray2 = mf_dot3(__ldg(k + 6), __ldg(k + 7), __ldg(k + 8), xf, yf);
    }

    const TIn *rgb_b = reinterpret_cast<const TIn *>(p.rgb) + (size_t)(b / p.poses) * p.rgb_sb;
    const TIn *sig_b = reinterpret_cast<const TIn *>(p.sigma) + (size_t)(b / p.poses) * p.sig_sb;
    const TIn *obj_b = HAS_OBJ ? reinterpret_cast<const TIn *>(p.obj) + (size_t)b * p.obj_sb : nullptr;

    TgtAcc t;
    float cnt = 0.f;
    // plane s-1 waits for plane s's xyz to know its thickness (:68-71)
    float pr = 0.f, pg = 0.f, pb = 0.f, psig = 0.f, pm = 0.f, px = 0.f, py = 0.f, pz = 0.f;
    // source frame
    // four partial flow sums (planes s % 4), like the persistent kernel: see planes_in_turns in render_pk.cuh
    float acc_s = 1.f, fu[4] = {0.f, 0.f, 0.f, 0.f}, fv[4] = {0.f, 0.f, 0.f, 0.f}, best_w = -1.f;
    const bool hard = (p.flags & MPIFLOW_HARD_FLOW) != 0;

    auto plane = [&](int s, auto acc_tag) {
        constexpr int A = decltype(acc_tag)::value;
        const PlaneConst &c = spc[s];
        // ---------------- target frame ----------------
        const SrcPos sp = tgt_to_src(c.hst, xf, yf, Wf, Hf, rW, rH);
        cnt += sp.valid ? 1.f : 0.f;  // :329-332, :347
        const float x0f = floorf(sp.ix), y0f = floorf(sp.iy);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);  // weight is 0 where clamped
        const float wx1 = __fsub_rn(__fadd_rn(x0f, 1.f), sp.ix), wx0 = __fsub_rn(sp.ix, x0f);
        const float wy1 = __fsub_rn(__fadd_rn(y0f, 1.f), sp.iy), wy0 = __fsub_rn(sp.iy, y0f);
        const float nw = __fmul_rn(wx1, wy1), ne = __fmul_rn(wx0, wy1);
        const float sw = __fmul_rn(wx1, wy0), se = __fmul_rn(wx0, wy0);
        const int o00 = y0 * W + x0, o01 = y0 * W + x1, o10 = y1 * W + x0, o11 = y1 * W + x1;
        auto tap = [&](const TIn *img) {
            float a = __fmul_rn(mf_load(img + o00), nw);
            a = __fmaf_rn(mf_load(img + o01), ne, a);
            a = __fmaf_rn(mf_load(img + o10), sw, a);
            return __fmaf_rn(mf_load(img + o11), se, a);
        };
        const TIn *rgb_s = rgb_b + (size_t)s * p.rgb_ss;
        const TIn *sig_s = sig_b + (size_t)s * p.sig_ss;
        const float cr = tap(rgb_s), cg = tap(rgb_s + p.rgb_sc), cb = tap(rgb_s + 2 * p.rgb_sc);
        float sg = tap(sig_s);
        const float m = HAS_OBJ ? tap(obj_b + (size_t)s * p.obj_ss) : 0.f;
        // plane xyz in the target frame at the clamped sample position
        const float X = fmaf(c.P[0], sp.ix, fmaf(c.P[1], sp.iy, c.P[2]));
        const float Y = fmaf(c.P[3], sp.ix, fmaf(c.P[4], sp.iy, c.P[5]));
        const float Z = fmaf(c.P[6], sp.ix, fmaf(c.P[7], sp.iy, c.P[8]));
        sg = (Z >= 0.f) ? sg : 0.f;  // :334-338
        if (s > 0) {
            const float dx = X - px, dy = Y - py, dz = Z - pz;
            t.add(psig, sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx))), pr, pg, pb, pz, pm);
        }
        pr = cr; pg = cg; pb = cb; psig = sg; pm = m; px = X; py = Y; pz = Z;

        // ---------------- source frame ----------------
        const float U = mf_dot3(c.g01[0], c.g01[1], c.g2[0], xf, yf);
        const float V = mf_dot3(c.g01[2], c.g01[3], c.g2[1], xf, yf);
        const float Q = mf_dot3(c.g01[4], c.g01[5], c.g2[2], xf, yf);
        float tu, tv;
        mf_div2(U, V, Q, tu, tv);
        const float du = __fsub_rn(tu, xf), dv = __fsub_rn(tv, yf);  // :212-218
        // thickness literally: |fl(ray * depth[s+1]) - fl(ray * depth[s])|; transmittance with the reference's expf
        const float ex = __fsub_rn(__fmul_rn(ray0, c.depth_next), __fmul_rn(ray0, c.depth));
        const float ey = __fsub_rn(__fmul_rn(ray1, c.depth_next), __fmul_rn(ray1, c.depth));
        const float ez = __fsub_rn(__fmul_rn(ray2, c.depth_next), __fmul_rn(ray2, c.depth));
        const float dist_s = (s + 1 < S) ? sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez))) : 1e3f;
        const float Ts = expf(__fmul_rn(-mf_load(sig_s + q), dist_s));
        const float ws = acc_s * (1.f - Ts);
        acc_s *= (Ts + 1e-6f);
        if (hard) {  // one-hot of the first arg-max weight (:126-130)
            if (ws > best_w) { best_w = ws; fu[0] = du; fv[0] = dv; }
        } else {
            fu[A] = fmaf(ws, du, fu[A]);
            fv[A] = fmaf(ws, dv, fv[A]);
        }
    };
    {
        int s = 0;
        for (; s + 4 <= S; s += 4) {
            plane(s, std::integral_constant<int, 0>{});
            plane(s + 1, std::integral_constant<int, 1>{});
            plane(s + 2, std::integral_constant<int, 2>{});
            plane(s + 3, std::integral_constant<int, 3>{});
        }
        for (; s < S; s++) plane(s, std::integral_constant<int, 0>{});
    }
    t.add(psig, 1e3f, pr, pg, pb, pz, pm);  // last plane: thickness 1e3 (:73-78)

    float *o = p.out_rgb + (size_t)b * 3 * HW + q;
    o[0] = t.r; o[HW] = t.g; o[2 * HW] = t.b;
    p.out_depth[(size_t)b * HW + q] = (p.flags & MPIFLOW_BG_DEPTH_INF)
                                          ? fmaf(1.f - t.wsum, 1000.f, t.zsum)      // :148-149
                                          : __fdiv_rn(t.zsum, t.wsum + 1e-5f);      // :151-152
    p.out_mask[(size_t)b * HW + q] = cnt;
    p.out_flow[(size_t)b * 2 * HW + q] = hard ? fu[0] : __fadd_rn(__fadd_rn(fu[0], fu[1]), __fadd_rn(fu[2], fu[3]));
    p.out_flow[(size_t)b * 2 * HW + HW + q] = hard ? fv[0] : __fadd_rn(__fadd_rn(fv[0], fv[1]), __fadd_rn(fv[2], fv[3]));
    if (p.out_obj) p.out_obj[(size_t)b * HW + q] = t.obj;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

MPIFLOW_API size_t mpiflow_render_workspace_bytes(int B, int S)
{
    if (B <= 0 || S <= 0) return 0;
    // one record more than B*S: the persistent kernel copies S + 1 records per tile (render_pk.cuh)
    return align_up(((size_t)B * S + 1) * sizeof(PlaneConst), 256) + align_up((size_t)B * sizeof(BatchConst), 256) + 256;
}

// the same plus room for the [S,H,W] plane-spacing table (see spacing_table_kernel): a launch that gets this much
// workspace evaluates the source-frame plane spacing once instead of per frame
MPIFLOW_API size_t mpiflow_render_workspace_bytes_hw(int B, int S, int H, int W)
{
    if (B <= 0 || S <= 0 || H <= 0 || W <= 0) return 0;
    return mpiflow_render_workspace_bytes(B, S) + align_up((size_t)S * H * W * sizeof(float), 256) +
           align_up((size_t)(TABLE_KEY_DEPTHS + S) * sizeof(uint32_t), 256);
}

// ---- TMA path ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// fp32 tensor map over (W, H, d2[, d3[, d4]]) with byte strides; box (bw, bh, b2, 1, 1)
static bool make_map(CUtensorMap *m, const float *ptr, int rank, const uint64_t *dims, const uint64_t *strides_b,
                     uint32_t bw, uint32_t bh, uint32_t b2)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    cuuint64_t gd[5], gs[4];
    cuuint32_t box[5] = {bw, bh, b2, 1, 1}, es[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; i++) gd[i] = dims[i];
    for (int i = 0; i < rank - 1; i++) {
        gs[i] = strides_b[i];
        if (gs[i] % 16 != 0 || gs[i] == 0) return false;
    }
    if (((uintptr_t)ptr & 15) != 0) return false;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void *)ptr, gd, gs, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tensor maps over the caller's planar tensors: rgb (W,H,3,S,B) box (bw,bh,3), sigma / mask (W,H,S,B) box
// (bw,bh), and sigma again with the tile itself as the box (un-warped sigma of the source-frame weights)
bool mf_make_render_maps(const RenderParams &p, uint32_t bw, uint32_t bh, uint32_t tw, uint32_t th, RenderMaps *m)
{
    const int B = p.B, S = p.S, H = p.H, W = p.W;
    if ((W % 4) != 0) return false;
    const uint64_t w4 = (uint64_t)W * 4, hw4 = (uint64_t)H * W * 4;
    {
        // a batch stride of 0 (one MPI under several poses) is a batch extent of 1; the kernels then ask for entry 0
        const bool one = B == p.poses || p.rgb_sb == 0;
        const uint64_t d[5] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)S, (uint64_t)(one ? 1 : B / p.poses)};
        const uint64_t sb[4] = {w4, (uint64_t)p.rgb_sc * 4, (uint64_t)(S > 1 ? p.rgb_ss : p.rgb_sc * 3) * 4,
                                (uint64_t)(one ? (int64_t)hw4 : p.rgb_sb * 4)};
        if (p.rgb_sc <= 0 || p.rgb_ss < 0 || p.rgb_sb < 0) return false;
        if (!make_map(&m->rgb, p.rgb, 5, d, sb, bw, bh, 3)) return false;
    }
    {
        const bool one = B == p.poses || p.sig_sb == 0;
        const uint64_t d[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)S, (uint64_t)(one ? 1 : B / p.poses)};
        const uint64_t sb[3] = {w4, (uint64_t)(S > 1 ? p.sig_ss * 4 : (int64_t)hw4), (uint64_t)(one ? (int64_t)hw4 : p.sig_sb * 4)};
        if (p.sig_ss < 0 || p.sig_sb < 0) return false;
        if (!make_map(&m->sig, p.sigma, 4, d, sb, bw, bh, 1)) return false;
        if (!make_map(&m->sigd, p.sigma, 4, d, sb, tw, th, 1)) return false;
    }
    if (p.obj) {
        const bool shared_mask = (p.obj_ss == 0) || S == 1;
        const bool one = B == 1 || p.obj_sb == 0;
        const uint64_t d[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)(shared_mask ? 1 : S), (uint64_t)(one ? 1 : B)};
        const uint64_t sb[3] = {w4, (uint64_t)(shared_mask ? (int64_t)hw4 : p.obj_ss * 4), (uint64_t)(one ? (int64_t)hw4 : p.obj_sb * 4)};
        if (p.obj_ss < 0 || p.obj_sb < 0) return false;
        if (!make_map(&m->obj, p.obj, 4, d, sb, bw, bh, 1)) return false;
    } else {
        m->obj = m->sig;
    }
    m->dist = m->sigd;
    if (p.spacing_table) {   // [S,H,W] table of batch entry 0, one tile per (tile, plane)
        const uint64_t d[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)S};
        const uint64_t sb[2] = {w4, hw4};
        if (!make_map(&m->dist, p.spacing_table, 3, d, sb, tw, th, 1)) return false;
    }
    return true;
}

int mf_sm_count()
{
    // per device: a process may render on several GPUs
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev];
    if (v > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    cached[dev] = v;
    return v;
}

// persistent kernel shapes (render_pk16.cu / render_pk8.cu).  Default: tile 64x16, box 80x20 (zoom up to ~1.19),
// 5-stage ring, one pixel pair per thread, 16+2 warps; 64x8 tiles with 8+2 warps for small launches (few tiles per
// SM).  The ring gets shorter when the plane constants of S planes (two tiles' worth) need the room.
static int pk_cfg()
{
    static const int v = []() {
        const char *e = getenv("MPIFLOW_PK_CFG");
        return e ? atoi(e) : 0;
    }();
    return v;
}

// MPIFLOW_RENDER_PATH = pk (default: persistent TMA kernel) | generic (plain global loads); debugging / A-B
// measurements only -- tensors the persistent kernel cannot take fall through to the generic one
static int render_path()
{
    static const int v = []() {
        const char *e = getenv("MPIFLOW_RENDER_PATH");
        if (e && strcmp(e, "generic") == 0) return 2;
        return 0;
    }();
    return v;
}

MPIFLOW_API int mpiflow_render_fused_f32(const float *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc,
                                         const float *sigma, int64_t sig_sb, int64_t sig_ss,
                                         const float *obj, int64_t obj_sb, int64_t obj_ss,
                                         const float *disparity, const float *G, const float *Kinv,
                                         const float *Kt, int B, int S, int H, int W, unsigned flags,
                                         float *out_rgb, float *out_depth, float *out_mask,
                                         float *out_flow, float *out_obj, void *workspace,
                                         size_t workspace_bytes, int *status, void *stream)
{
    return mf_render_fused_impl(rgb, rgb_sb, rgb_ss, rgb_sc, sigma, sig_sb, sig_ss, obj, obj_sb, obj_ss, disparity, G, Kinv,
                                Kt, B, S, H, W, flags, out_rgb, out_depth, out_mask, out_flow, out_obj, workspace,
                                workspace_bytes, status, /*clear_status=*/true, /*f16=*/false, (cudaStream_t)stream);
}

// fp16 storage: rgb / sigma / obj are IEEE halves (strides in elements), everything else as above; fp32 arithmetic on
// the converted values.  Served by the plain global-load kernel (half the HBM bytes per tap, no shared-memory staging):
// it is the entry for callers whose bottleneck is moving the MPI, e.g. over PCIe (mpiflow_render_fused_host_f16).
MPIFLOW_API int mpiflow_render_fused_f16(const void *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc,
                                         const void *sigma, int64_t sig_sb, int64_t sig_ss,
                                         const void *obj, int64_t obj_sb, int64_t obj_ss,
                                         const float *disparity, const float *G, const float *Kinv,
                                         const float *Kt, int B, int S, int H, int W, unsigned flags,
                                         float *out_rgb, float *out_depth, float *out_mask,
                                         float *out_flow, float *out_obj, void *workspace,
                                         size_t workspace_bytes, int *status, void *stream)
{
    return mf_render_fused_impl((const float *)rgb, rgb_sb, rgb_ss, rgb_sc, (const float *)sigma, sig_sb, sig_ss,
                                (const float *)obj, obj_sb, obj_ss, disparity, G, Kinv, Kt, B, S, H, W, flags, out_rgb,
                                out_depth, out_mask, out_flow, out_obj, workspace, workspace_bytes, status,
                                /*clear_status=*/true, /*f16=*/true, (cudaStream_t)stream);
}

// the same with a choice about the status word: the host-buffer entry point lets the frames of one stream share a
// sticky flag (cleared once per call) and reads it back once
int mf_render_fused_impl(const float *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc,
                         const float *sigma, int64_t sig_sb, int64_t sig_ss,
                         const float *obj, int64_t obj_sb, int64_t obj_ss,
                         const float *disparity, const float *G, const float *Kinv,
                         const float *Kt, int B, int S, int H, int W, unsigned flags,
                         float *out_rgb, float *out_depth, float *out_mask,
                         float *out_flow, float *out_obj, void *workspace,
                         size_t workspace_bytes, int *status, bool clear_status, bool f16, cudaStream_t st)
{
    MF_CHECK_ARG(rgb && sigma && disparity && G && Kinv && Kt && workspace);
    MF_CHECK_ARG(out_rgb && out_depth && out_mask && out_flow);
    MF_CHECK_ARG(B > 0 && S > 0 && H > 0 && W > 0 && B <= 65535 && (int64_t)H * W < (1ll << 31));
    MF_CHECK_ARG((flags & ~(MPIFLOW_BG_DEPTH_INF | MPIFLOW_HARD_FLOW | MPIFLOW_KEEP_TABLE | 0xffff00u)) == 0);
    const int poses = (int)((flags >> 8) & 0xffu) + 1;
    const int reserve_sms = (int)((flags >> 16) & 0xffu);
    MF_CHECK_ARG(B % poses == 0);
    const bool keep_table = (flags & MPIFLOW_KEEP_TABLE) != 0;
    flags &= ~(0xffff00u | MPIFLOW_KEEP_TABLE);
    MF_CHECK_ARG(((uintptr_t)workspace & 15) == 0);
    if (workspace_bytes < mpiflow_render_workspace_bytes(B, S)) return MPIFLOW_E_WORKSPACE;
    const size_t smem = (size_t)S * sizeof(PlaneConst);
    MF_CHECK_ARG(smem <= 200 * 1024);
    if (status && clear_status) {
        cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int), st);
        if (e != cudaSuccess) return (int)e;
    }
    PlaneConst *pc = reinterpret_cast<PlaneConst *>(workspace);
    BatchConst *bc = reinterpret_cast<BatchConst *>(reinterpret_cast<char *>(workspace) +
                                                    align_up(((size_t)B * S + 1) * sizeof(PlaneConst), 256));
    int *tile_counter = reinterpret_cast<int *>(reinterpret_cast<char *>(bc) + align_up((size_t)B * sizeof(BatchConst), 256));
    plane_consts_kernel<<<(B * S + 127) / 128, 128, 0, st>>>(G, Kinv, Kt, disparity, B, S, pc, bc, status, tile_counter);
    int rc = mf_launch_status();
    if (rc) return rc;
    // Source-frame plane spacing once per launch instead of per frame and plane (spacing_table_kernel): when the caller
    // gave room for the table (mpiflow_render_workspace_bytes_hw), the launch is large enough to amortise it, and TMA
    // can describe it.  Entries whose K_src^-1 or plane depths differ from entry 0 evaluate it in the plane loop.
    static const bool table_off = []() { const char *e = getenv("MPIFLOW_PK_SPACING_TABLE"); return e && atoi(e) == 0; }();
    const size_t base_bytes = mpiflow_render_workspace_bytes(B, S);
    const bool have_table = !f16 && !table_off && render_path() == 0 && (B >= 4 || keep_table) && (W % 4) == 0 &&
                            workspace_bytes >= mpiflow_render_workspace_bytes_hw(B, S, H, W);
    float *table = have_table ? reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + base_bytes) : nullptr;
    uint32_t *table_key = have_table ? reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(table) +
                                                                    align_up((size_t)S * H * W * sizeof(float), 256)) : nullptr;
    batch_flags_kernel<<<B + (have_table ? 1 : 0), 32, 0, st>>>(pc, bc, B, S, H, W, have_table ? 1 : 0, keep_table ? 1 : 0, table_key);
    rc = mf_launch_status();
    if (rc) return rc;
    if (have_table) {
        spacing_table_kernel<<<dim3((W / 4 + 127) / 128, H, (S + TABLE_PLANES - 1) / TABLE_PLANES), 128, 0, st>>>(pc, bc, S, H, W, table, table_key);
        rc = mf_launch_status();
        if (rc) return rc;
    }

    RenderParams p;
    p.rgb = rgb; p.rgb_sb = rgb_sb; p.rgb_ss = rgb_ss; p.rgb_sc = rgb_sc;
    p.sigma = sigma; p.sig_sb = sig_sb; p.sig_ss = sig_ss;
    p.obj = obj; p.obj_sb = obj_sb; p.obj_ss = obj_ss;
    p.pc = pc; p.bc = bc;
    p.B = B; p.S = S; p.H = H; p.W = W; p.flags = flags; p.poses = poses;
    p.out_rgb = out_rgb; p.out_depth = out_depth; p.out_mask = out_mask; p.out_flow = out_flow;
    p.out_obj = out_obj;
    p.spacing_table = table;
    p.reserve_sms = reserve_sms;
    p.hW = (float)W * 0.5f; p.hH = (float)H * 0.5f;
    p.rW = 1.0f / p.hW; p.rH = 1.0f / p.hH;    // IEEE quotients = __frcp_rn: "tensor / python scalar" on CUDA multiplies by this

    if (render_path() == 0 && !f16) {
        bool used = false;
        static const bool force_literal = []() { const char *e = getenv("MPIFLOW_PK_FORCE_LITERAL"); return e && atoi(e) != 0; }();
        if (force_literal) p.flags |= MF_FLAG_FORCE_LITERAL;
        static const unsigned dbg_bits = []() { const char *e = getenv("MPIFLOW_PK_DBG"); return e ? (unsigned)strtoul(e, nullptr, 0) : 0u; }();
        p.flags |= dbg_bits & (MF_FLAG_DBG_NO_LOADS | MF_FLAG_DBG_NO_COMPUTE);
        // few 64x16 tiles per SM (small images / batches): halve the tile so that the persistent CTAs stay balanced
        const long long tiles16 = (long long)((W + 63) / 64) * ((H + 15) / 16) * B;
        const bool small = tiles16 < 6ll * mf_sm_count();
        switch (pk_cfg()) {
            case 1: case 2: rc = mf_launch_pk_ab(pk_cfg(), p, tile_counter, st, &used); break;
            case 3: rc = mf_launch_pk8(8, p, tile_counter, st, &used); break;
            case 9: rc = mf_launch_pk16(5, p, tile_counter, st, &used); break;
            case 4: rc = mf_launch_pk16(15, p, tile_counter, st, &used);      // split warps, A/B
                    if (!rc && !used) rc = mf_launch_pk16(14, p, tile_counter, st, &used);
                    break;
            default: break;
        }
        // the deepest ring that leaves room for the plane constants of S planes
        if (!rc && !used && small) rc = mf_launch_pk8(8, p, tile_counter, st, &used);
        if (!rc && !used && small) rc = mf_launch_pk8(4, p, tile_counter, st, &used);
        if (!rc && !used) rc = mf_launch_pk16(5, p, tile_counter, st, &used);
        if (!rc && !used) rc = mf_launch_pk16(4, p, tile_counter, st, &used);
        if (rc || used) return rc;
    }
    // generic path: tensors TMA cannot describe (W % 4 != 0, unaligned base, odd strides)
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    auto kern = f16 ? (obj ? render_fused_kernel<true, __half> : render_fused_kernel<false, __half>)
                    : (obj ? render_fused_kernel<true, float> : render_fused_kernel<false, float>);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<grid, 256, smem, st>>>(p);
    return mf_launch_status();
}

// test hook: the shared-reciprocal division used in the kernels, element-wise on device arrays
__global__ void debug_div_kernel(const float *a, const float *c, const float *b, int n, float *qa, float *qc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mf_div2(a[i], c[i], b[i], qa[i], qc[i]);
}

MPIFLOW_API int mpiflow_debug_div_f32(const float *a, const float *c, const float *b, int n, float *qa, float *qc,
                                      void *stream)
{
    MF_CHECK_ARG(a && c && b && qa && qc && n > 0);
    debug_div_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, c, b, n, qa, qc);
    return mf_launch_status();
}

// test hook: the packed (f32x2) arithmetic of render_pk.cuh against the scalar IEEE operations, element-wise on
// device arrays of even length n; out[k*n + i], k = 0..9: fma2, fma, mul2, mul, sub2, sub, floor-by-magic, floorf,
// unfused a*b+c packed, scalar, packed expf clone of -|a|/22, expf of the same, packed sqrt of |a*b|, sqrtf
__global__ void debug_packed_kernel(const float *a, const float *b, const float *c, int n, float *out)
{
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 1 >= n) return;
    const float2 A = make_float2(a[i], a[i + 1]), B = make_float2(b[i], b[i + 1]), C = make_float2(c[i], c[i + 1]);
    const float2 f = __ffma2_rn(A, B, C), m = __fmul2_rn(A, pk(B.x)), d = __fadd2_rn(A, pkneg(C));
    out[0 * n + i] = f.x; out[0 * n + i + 1] = f.y;
    out[1 * n + i] = __fmaf_rn(A.x, B.x, C.x); out[1 * n + i + 1] = __fmaf_rn(A.y, B.y, C.y);
    out[2 * n + i] = m.x; out[2 * n + i + 1] = m.y;
    out[3 * n + i] = __fmul_rn(A.x, B.x); out[3 * n + i + 1] = __fmul_rn(A.y, B.x);
    out[4 * n + i] = d.x; out[4 * n + i + 1] = d.y;
    out[5 * n + i] = __fsub_rn(A.x, C.x); out[5 * n + i + 1] = __fsub_rn(A.y, C.y);
    const float2 pos = make_float2(fabsf(A.x), fabsf(A.y));
    const float2 t = __fadd2_rd(pos, pk(MF_MAGIC));
    out[6 * n + i] = (float)(__float_as_uint(t.x) - MF_MAGIC_BITS); out[6 * n + i + 1] = (float)(__float_as_uint(t.y) - MF_MAGIC_BITS);
    out[7 * n + i] = floorf(pos.x); out[7 * n + i + 1] = floorf(pos.y);
    const float2 u = pk_mul_then_add(A, pk(B.x), pk(-1.f));
    out[8 * n + i] = u.x; out[8 * n + i + 1] = u.y;
    out[9 * n + i] = __fadd_rn(__fmul_rn(A.x, B.x), -1.f); out[9 * n + i + 1] = __fadd_rn(__fmul_rn(A.y, B.x), -1.f);
    const float2 xe = make_float2(-fabsf(A.x) * (1.f / 22.f), -fabsf(A.y) * (1.f / 22.f));   // down to -95
    const float2 ee = pk_expf(xe);
    out[10 * n + i] = ee.x; out[10 * n + i + 1] = ee.y;
    out[11 * n + i] = expf(xe.x); out[11 * n + i + 1] = expf(xe.y);
    const float2 qa = make_float2(fabsf(A.x * B.x), fabsf(A.y * B.y));
    const float2 qs = pk_sqrt_rn(qa);
    out[12 * n + i] = qs.x; out[12 * n + i + 1] = qs.y;
    out[13 * n + i] = sqrtf(fmaxf(qa.x, 1e-30f)); out[13 * n + i + 1] = sqrtf(fmaxf(qa.y, 1e-30f));
}

MPIFLOW_API int mpiflow_debug_packed_f32(const float *a, const float *b, const float *c, int n, float *out, void *stream)
{
    MF_CHECK_ARG(a && b && c && out && n > 0 && (n % 2) == 0);
    debug_packed_kernel<<<(n / 2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, b, c, n, out);
    return mf_launch_status();
}
