/*
 * oracle/mpi_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32 unless stated) of the MPI-Flow renderer hot path, written
 * from the reference's behaviour, function by function, so that the sm_100a kernels can be
 * checked against it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.  Nothing under mpi-flow_b200/ may.
 *
 * Parity pin: the reference ships no tests or golden vectors for this path,
 * so this oracle is pinned against outputs of the reference itself, run in the build
 * container on CPU (tests/golden/make_golden.py -> the .npz files in tests/golden/, checked by
 * tests/test_oracle_golden.py).
 *
 * Each function cites the reference lines it restates (paths relative to the reference root).
 *
 * "numerics" selects which library arithmetic is mimicked where torch-CPU and torch-CUDA differ
 * by an ulp (all four differences are documented in DESIGN.md section "numerics contract"):
 *   ORC_NUM_CPU  (0): 3x3@3x3 products without FMA (ATen small-bmm path), true division by the
 *                     python scalar W*0.5, grid_sample unnormalise fma(g+1, W/2, -0.5)
 *   ORC_NUM_CUDA (1): FMA chains (cuBLAS), multiply by the reciprocal of the scalar
 *                     (ATen BinaryDivTrueKernel.cu), unnormalise fma(g+1, W, -1)*0.5
 *                     (ATen GridSampler.cuh:29 compiled with nvcc's default -fmad=true)
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -mfma -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NUM_CPU 0
#define ORC_NUM_CUDA 1

/* ------------------------------------------------------------------------------------------ */
/* small dense helpers                                                                         */
/* ------------------------------------------------------------------------------------------ */

/* k-ordered dot product the way the library GEMM behind torch.matmul accumulates it.
 * fused=1: acc = a0*b0; acc = fma(a1,b1,acc); ...   (MKL sgemm / cuBLAS; probed bit-exact on
 * torch-CPU 2.11 for [B,3,3]x[B,3,HW]);  fused=0: ((a0*b0 + a1*b1) + a2*b2) with each product
 * rounded (ATen baddbmm_cpu_kernel, used when M*N*K < 400, i.e. the 3x3@3x3 products). */
static inline float dotk(const float *a, int sa, const float *b, int sb, int k, int fused)
{
    float acc = a[0] * b[0];
    for (int i = 1; i < k; i++) {
        if (fused)
            acc = fmaf(a[i * sa], b[i * sb], acc);
        else
            acc = acc + a[i * sa] * b[i * sb];
    }
    return acc;
}

static void matmul33(const float *A, const float *B, float *C, int fused)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C[i * 3 + j] = dotk(A + i * 3, 1, B + j, 3, 3, fused);
}

/* fp64 3x3 inverse, Gauss-Jordan with partial pivoting (stands in for LAPACK getrf/getri behind
 * torch.inverse on the CPU, utils/mpi/homography_sampler.py:6-27,121-122). Returns 0 on success. */
static int inverse33_f64(const double *A, double *Ainv)
{
    double m[3][6];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            m[i][j] = A[i * 3 + j];
            m[i][3 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 3; c++) {
        int p = c;
        for (int r = c + 1; r < 3; r++)
            if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (m[p][c] == 0.0) return 1;
        if (p != c)
            for (int j = 0; j < 6; j++) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
        double inv = 1.0 / m[c][c];
        for (int j = 0; j < 6; j++) m[c][j] *= inv;
        for (int r = 0; r < 3; r++) {
            if (r == c) continue;
            double f = m[r][c];
            for (int j = 0; j < 6; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Ainv[i * 3 + j] = m[i][3 + j];
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* a4/a5 prologue: plane-induced homographies                                                  */
/* utils/mpi/homography_sampler.py:96-122 (sample) and :176-197 (sample_inverse)               */
/* ------------------------------------------------------------------------------------------ */

/* N = number of (batch*plane) entries; G [N,4,4], Kinv/Kt [N,3,3], depth [N] (= 1/disparity).
 * Writes H_tgt_src and H_src_tgt, both [N,9] fp32.  Returns the number of singular matrices. */
int orc_plane_homographies(const float *G, const float *Kinv, const float *Kt, const float *depth,
                           int N, int numerics, float *Hts, float *Hst)
{
    int bad = 0;
    const int fused = (numerics == ORC_NUM_CUDA);
    for (int n = 0; n < N; n++) {
        const float *g = G + n * 16;
        const float nvec[3] = {0.f, 0.f, 1.f}; /* plane normal, :58-61 */
        float Rtnd[9], M1[9];
        const float negd = -depth[n];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                float tn = g[i * 4 + 3] * nvec[j];         /* matmul(t[3x1], n[1x3]), k = 1 */
                Rtnd[i * 3 + j] = g[i * 4 + j] - tn / negd; /* :114-116, "/ -d" */
            }
        matmul33(Rtnd, Kinv + n * 9, M1, fused);           /* :117-118 */
        matmul33(Kt + n * 9, M1, Hts + n * 9, fused);
        double h64[9], hi64[9];
        for (int i = 0; i < 9; i++) h64[i] = (double)Hts[n * 9 + i];
        if (inverse33_f64(h64, hi64)) { bad++; for (int i = 0; i < 9; i++) hi64[i] = NAN; }
        for (int i = 0; i < 9; i++) Hst[n * 9 + i] = (float)hi64[i]; /* :121-122 */
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------ */
/* a2/a3: plane xyz in source and target camera frames                                         */
/* ------------------------------------------------------------------------------------------ */

/* utils/mpi/mpi_rendering.py:213-239.  Kinv [B,3,3], disparity [B,S] -> xyz [B,S,3,H,W].
 * meshgrid is the integer pixel grid (x, y, 1) of homography_sampler.py:46-56. */
void orc_src_xyz_from_plane_disparity(const float *Kinv, const float *disparity, int B, int S,
                                      int H, int W, float *xyz)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; b++)
        for (int s = 0; s < S; s++) {
            const float *k = Kinv + b * 9;
            const float depth = 1.0f / disparity[b * S + s]; /* torch.reciprocal, :224 */
            float *o = xyz + ((size_t)(b * S + s)) * 3 * HW;
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    const float p[3] = {(float)x, (float)y, 1.f};
                    for (int i = 0; i < 3; i++) {
                        float v = dotk(k + i * 3, 1, p, 1, 3, 1); /* matmul [3x3]x[3xHW], :234 */
                        o[i * HW + (size_t)y * W + x] = v * depth;  /* :235-236 */
                    }
                }
        }
}

/* utils/mpi/mpi_rendering.py:242-256 + utils/mpi/rendering_utils.py:4-23.
 * G [B,4,4]; xyz_src [B,S,3,H,W] -> xyz_tgt [B,S,3,H,W] = (G . [xyz;1])[:3]. */
void orc_tgt_xyz_from_plane_disparity(const float *xyz_src, const float *G, int B, int S, int H,
                                      int W, float *xyz_tgt)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; b++)
        for (int s = 0; s < S; s++) {
            const float *g = G + b * 16;
            const float *in = xyz_src + ((size_t)(b * S + s)) * 3 * HW;
            float *o = xyz_tgt + ((size_t)(b * S + s)) * 3 * HW;
            for (size_t p = 0; p < HW; p++) {
                const float v[4] = {in[p], in[HW + p], in[2 * HW + p], 1.f};
                for (int i = 0; i < 3; i++) o[i * HW + p] = dotk(g + i * 4, 1, v, 1, 4, 1);
            }
        }
}

/* ------------------------------------------------------------------------------------------ */
/* a4: HomographySample.sample  (utils/mpi/homography_sampler.py:80-158)                        */
/* ------------------------------------------------------------------------------------------ */

/* target pixel (x,y) -> source position, validity and grid_sample's clipped source index */
static inline void tgt_to_src(const float *h, float x, float y, int H, int W, int numerics,
                              float *sx_o, float *sy_o, int *valid, float *ix_o, float *iy_o)
{
    const float p[3] = {x, y, 1.f};
    const float X = dotk(h + 0, 1, p, 1, 3, 1); /* :131-132 */
    const float Y = dotk(h + 3, 1, p, 1, 3, 1);
    const float Z = dotk(h + 6, 1, p, 1, 3, 1);
    const float sx = X / Z, sy = Y / Z;         /* :137-138 */
    *sx_o = sx;
    *sy_o = sy;
    *valid = (sx < (float)W) && (sx > -1.f) && (sy < (float)H) && (sy > -1.f); /* :143-147 */
    float gx, gy, ix, iy;
    if (numerics == ORC_NUM_CUDA) {
        /* tensor / python-scalar on CUDA multiplies by the fp32 reciprocal */
        const float rw = 1.0f / ((float)W * 0.5f), rh = 1.0f / ((float)H * 0.5f);
        gx = (sx + 0.5f) * rw - 1.f; /* :151-154 */
        gy = (sy + 0.5f) * rh - 1.f;
        ix = fmaf(gx + 1.f, (float)W, -1.f) * 0.5f; /* GridSampler.cuh:29, nvcc -fmad */
        iy = fmaf(gy + 1.f, (float)H, -1.f) * 0.5f;
    } else {
        gx = (sx + 0.5f) / ((float)W * 0.5f) - 1.f;
        gy = (sy + 0.5f) / ((float)H * 0.5f) - 1.f;
        /* ATen cpu/GridSamplerKernel.cpp unnormalize: (g+1)*(size/2) - 0.5, contracted to one FMA
         * in the shipped AVX2 build (probed: the fused form reproduces torch-CPU 2.11 to 1e-7, the
         * unfused one only to 1e-5 on white-noise inputs) */
        ix = fmaf(gx + 1.f, (float)W / 2.f, -0.5f);
        iy = fmaf(gy + 1.f, (float)H / 2.f, -0.5f);
    }
    /* padding_mode='border': clip to [0, size-1] (GridSampler.cuh:55-56) */
    *ix_o = fminf((float)(W - 1), fmaxf(ix, 0.f));
    *iy_o = fminf((float)(H - 1), fmaxf(iy, 0.f));
}

/* bilinear tap of one channel plane at a clipped index (ATen grid_sampler_2d, bilinear) */
static inline float bilinear_tap(const float *img, int H, int W, float ix, float iy)
{
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    const float nw = ((x0f + 1.f) - ix) * ((y0f + 1.f) - iy);
    const float ne = (ix - x0f) * ((y0f + 1.f) - iy);
    const float sw = ((x0f + 1.f) - ix) * (iy - y0f);
    const float se = (ix - x0f) * (iy - y0f);
    float acc = 0.f; /* out_acc += value * weight for each in-bounds corner, nw ne sw se */
    acc = fmaf(img[(size_t)y0 * W + x0], nw, acc);
    if (x1 <= W - 1) acc = fmaf(img[(size_t)y0 * W + x1], ne, acc);
    if (y1 <= H - 1) acc = fmaf(img[(size_t)y1 * W + x0], sw, acc);
    if (x1 <= W - 1 && y1 <= H - 1) acc = fmaf(img[(size_t)y1 * W + x1], se, acc);
    return acc;
}

/* src [N,C,H,W], Hst [N,9] -> tgt [N,C,H,W], valid [N,H,W] (u8), flowB2A [N,H,W,2].
 * Source and target sizes are equal (as in every reference call site). */
void orc_homography_sample(const float *src, const float *Hst, int N, int C, int H, int W,
                           int numerics, float *tgt, uint8_t *valid, float *flowB2A)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                float sx, sy, ix, iy;
                int v;
                tgt_to_src(Hst + n * 9, (float)x, (float)y, H, W, numerics, &sx, &sy, &v, &ix, &iy);
                const size_t p = (size_t)y * W + x;
                if (valid) valid[n * HW + p] = (uint8_t)v;
                if (flowB2A) {
                    flowB2A[(n * HW + p) * 2 + 0] = sx - (float)x; /* :139-141 */
                    flowB2A[(n * HW + p) * 2 + 1] = sy - (float)y;
                }
                for (int c = 0; c < C; c++)
                    tgt[((size_t)n * C + c) * HW + p] =
                        bilinear_tap(src + ((size_t)n * C + c) * HW, H, W, ix, iy); /* :155 */
            }
}

/* a5: HomographySample.sample_inverse (utils/mpi/homography_sampler.py:160-220).
 * Hts [N,9] -> flowA2B [N,H,W,2] = (H_ts . p)/Z - p for every SOURCE pixel p. */
void orc_homography_sample_inverse(const float *Hts, int N, int H, int W, float *flowA2B)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const float *h = Hts + n * 9;
                const float p[3] = {(float)x, (float)y, 1.f};
                const float X = dotk(h + 0, 1, p, 1, 3, 1); /* :207-208 */
                const float Y = dotk(h + 3, 1, p, 1, 3, 1);
                const float Z = dotk(h + 6, 1, p, 1, 3, 1);
                const size_t q = (n * HW + (size_t)y * W + x) * 2;
                flowA2B[q + 0] = X / Z - (float)x; /* :212-218 */
                flowA2B[q + 1] = Y / Z - (float)y;
            }
}

/* ------------------------------------------------------------------------------------------ */
/* a7/a8/a9: plane volume rendering                                                            */
/* ------------------------------------------------------------------------------------------ */

/* utils/mpi/mpi_rendering.py:62-99 (+ weighted_sum_mpi :142-154).
 * rgb [B,S,3,H,W], sigma [B,S,1,H,W], xyz [B,S,3,H,W]; optional src_flow [B,S,2,H,W] and
 * obj [B,S,1,H,W] (both or neither, as at the reference's call sites).
 * Outputs (any may be NULL): rgb_out [B,3,H,W], depth_out [B,1,H,W], transparency_acc
 * [B,S,1,H,W], weights [B,S,1,H,W], flow_out [B,2,H,W], obj_out [B,1,H,W]. */
void orc_plane_volume_rendering(const float *rgb, const float *sigma, const float *xyz,
                                const float *src_flow, const float *obj, int B, int S, int H,
                                int W, int is_bg_depth_inf, float *rgb_out, float *depth_out,
                                float *transparency_acc, float *weights, float *flow_out,
                                float *obj_out)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; b++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const size_t p = (size_t)y * W + x;
                float acc = 1.f; /* cumprod shifted by one, :84-87 */
                float wsum = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, zsum = 0.f;
                float f0 = 0.f, f1 = 0.f, osum = 0.f;
                for (int s = 0; s < S; s++) {
                    const size_t bs = (size_t)b * S + s;
                    const float *q = xyz + bs * 3 * HW + p;
                    float dist;
                    if (s < S - 1) { /* :68-71 */
                        const float dx = q[3 * HW] - q[0];
                        const float dy = q[3 * HW + HW] - q[HW];
                        const float dz = q[3 * HW + 2 * HW] - q[2 * HW];
                        dist = sqrtf(dx * dx + dy * dy + dz * dz);
                    } else {
                        dist = 1e3f; /* :73-78 */
                    }
                    const float T = expf(-sigma[bs * HW + p] * dist); /* :79 */
                    const float alpha = 1.f - T;                       /* :80 */
                    const float w = acc * alpha;                       /* :89 */
                    if (transparency_acc) transparency_acc[bs * HW + p] = acc;
                    if (weights) weights[bs * HW + p] = w;
                    acc = acc * (T + 1e-6f); /* :84 */
                    wsum += w;               /* :143 */
                    c0 += w * rgb[bs * 3 * HW + p];
                    c1 += w * rgb[bs * 3 * HW + HW + p];
                    c2 += w * rgb[bs * 3 * HW + 2 * HW + p];
                    zsum += w * q[2 * HW];
                    if (src_flow) { /* :92-96 */
                        f0 += w * src_flow[bs * 2 * HW + p];
                        f1 += w * src_flow[bs * 2 * HW + HW + p];
                        osum += w * obj[bs * HW + p];
                    }
                }
                if (rgb_out) {
                    rgb_out[((size_t)b * 3 + 0) * HW + p] = c0;
                    rgb_out[((size_t)b * 3 + 1) * HW + p] = c1;
                    rgb_out[((size_t)b * 3 + 2) * HW + p] = c2;
                }
                if (depth_out)
                    depth_out[(size_t)b * HW + p] = is_bg_depth_inf
                                                        ? zsum + (1.f - wsum) * 1000.f /* :148-149 */
                                                        : zsum / (wsum + 1e-5f);       /* :151-152 */
                if (src_flow && flow_out) {
                    flow_out[((size_t)b * 2 + 0) * HW + p] = f0;
                    flow_out[((size_t)b * 2 + 1) * HW + p] = f1;
                }
                if (src_flow && obj_out) obj_out[(size_t)b * HW + p] = osum;
            }
}

/* utils/mpi/mpi_rendering.py:102-139: same weights, built from the UN-warped sigma/xyz, applied to
 * the per-plane forward flow.  hard_flow: one-hot of the first arg-max weight (:126-130). */
void orc_plane_volume_rendering_flow(const float *src_sigma, const float *src_flow,
                                     const float *src_xyz, int B, int S, int H, int W,
                                     int hard_flow, float *flow_out)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; b++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const size_t p = (size_t)y * W + x;
                float acc = 1.f, f0 = 0.f, f1 = 0.f, best = -INFINITY;
                for (int s = 0; s < S; s++) {
                    const size_t bs = (size_t)b * S + s;
                    const float *q = src_xyz + bs * 3 * HW + p;
                    float dist;
                    if (s < S - 1) {
                        const float dx = q[3 * HW] - q[0];
                        const float dy = q[3 * HW + HW] - q[HW];
                        const float dz = q[3 * HW + 2 * HW] - q[2 * HW];
                        dist = sqrtf(dx * dx + dy * dy + dz * dz);
                    } else {
                        dist = 1e3f;
                    }
                    const float T = expf(-src_sigma[bs * HW + p] * dist);
                    const float w = acc * (1.f - T);
                    acc = acc * (T + 1e-6f);
                    const float u = src_flow[bs * 2 * HW + p], v = src_flow[bs * 2 * HW + HW + p];
                    if (hard_flow) {
                        if (w > best) { best = w; f0 = u; f1 = v; }
                    } else {
                        f0 += w * u;
                        f1 += w * v;
                    }
                }
                flow_out[((size_t)b * 2 + 0) * HW + p] = f0;
                flow_out[((size_t)b * 2 + 1) * HW + p] = f1;
            }
}

/* utils/mpi/mpi_rendering.py:42-59.  alpha [B,K,1,H,W], value [B,K,C,H,W] ->
 * composed [B,C,H,W], weights [B,K,1,H,W] (may be NULL). */
void orc_alpha_composition(const float *alpha, const float *value, int B, int K, int C, int H,
                           int W, float *composed, float *weights)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; b++)
        for (size_t p = 0; p < HW; p++) {
            float preserve = 1.f;
            float accum[16];
            for (int c = 0; c < C; c++) accum[c] = 0.f;
            for (int k = 0; k < K; k++) {
                const size_t bk = (size_t)b * K + k;
                const float a = alpha[bk * HW + p];
                const float w = a * preserve; /* :56 */
                if (weights) weights[bk * HW + p] = w;
                for (int c = 0; c < C; c++) accum[c] += value[(bk * C + c) * HW + p] * w;
                preserve = preserve * (1.f - a); /* cumprod(1 - alpha), :52 */
            }
            for (int c = 0; c < C; c++) composed[((size_t)b * C + c) * HW + p] = accum[c];
        }
}

/* ------------------------------------------------------------------------------------------ */
/* a10: render_tgt_rgb_depth  (utils/mpi/mpi_rendering.py:259-349), sigma mode                  */
/* ------------------------------------------------------------------------------------------ */

/* Literal orchestration with materialised temporaries, like the reference:
 * cat(rgb, sigma, xyz_tgt[, obj]) -> sample -> split -> sigma=0 where z<0 -> render.
 * mpi_rgb [B,S,3,H,W], mpi_sigma [B,S,1,H,W], disparity [B,S], xyz_tgt/xyz_src [B,S,3,H,W],
 * G [B,4,4], Kinv/Kt [B,3,3], obj [B,S,1,H,W] (required, as utils/utils.py:328-346 always
 * passes it; the reference crashes at :95 without it).
 * Outputs: rgb [B,3,H,W], depth [B,1,H,W], mask(count) [B,1,H,W], flowA2B [B,2,H,W],
 * obj_out [B,1,H,W].  Returns 0, or -1 on allocation failure, or >0 = singular homographies. */
int orc_render_tgt_rgb_depth(const float *mpi_rgb, const float *mpi_sigma, const float *disparity,
                             const float *xyz_tgt, const float *xyz_src, const float *G,
                             const float *Kinv, const float *Kt, const float *obj, int B, int S,
                             int H, int W, int is_bg_depth_inf, int hard_flow, int numerics,
                             float *out_rgb, float *out_depth, float *out_mask, float *out_flow,
                             float *out_obj)
{
    const size_t HW = (size_t)H * W;
    const int N = B * S, C = 8;
    float *depth = malloc(sizeof(float) * N), *Gn = malloc(sizeof(float) * 16 * N);
    float *Kin = malloc(sizeof(float) * 9 * N), *Ktn = malloc(sizeof(float) * 9 * N);
    float *Hts = malloc(sizeof(float) * 9 * N), *Hst = malloc(sizeof(float) * 9 * N);
    float *cat = malloc(sizeof(float) * N * C * HW), *warped = malloc(sizeof(float) * N * C * HW);
    uint8_t *valid = malloc(N * HW);
    float *flow_planes = malloc(sizeof(float) * N * 2 * HW);  /* [N,H,W,2] then [N,2,H,W] */
    float *flow_chw = malloc(sizeof(float) * N * 2 * HW);
    float *t_rgb = malloc(sizeof(float) * N * 3 * HW), *t_sig = malloc(sizeof(float) * N * HW);
    float *t_xyz = malloc(sizeof(float) * N * 3 * HW), *t_obj = malloc(sizeof(float) * N * HW);
    int rc = -1;
    if (!depth || !Gn || !Kin || !Ktn || !Hts || !Hst || !cat || !warped || !valid ||
        !flow_planes || !flow_chw || !t_rgb || !t_sig || !t_xyz || !t_obj)
        goto done;

    for (int b = 0; b < B; b++)
        for (int s = 0; s < S; s++) {
            const int n = b * S + s;
            depth[n] = 1.0f / disparity[n]; /* :278 */
            memcpy(Gn + n * 16, G + b * 16, sizeof(float) * 16); /* :286-291 repeat over S */
            memcpy(Kin + n * 9, Kinv + b * 9, sizeof(float) * 9);
            memcpy(Ktn + n * 9, Kt + b * 9, sizeof(float) * 9);
            float *c = cat + (size_t)n * C * HW; /* :282-283, :297-298 */
            memcpy(c, mpi_rgb + (size_t)n * 3 * HW, sizeof(float) * 3 * HW);
            memcpy(c + 3 * HW, mpi_sigma + (size_t)n * HW, sizeof(float) * HW);
            memcpy(c + 4 * HW, xyz_tgt + (size_t)n * 3 * HW, sizeof(float) * 3 * HW);
            memcpy(c + 7 * HW, obj + (size_t)n * HW, sizeof(float) * HW);
        }
    rc = orc_plane_homographies(Gn, Kin, Ktn, depth, N, numerics, Hts, Hst);
    orc_homography_sample(cat, Hst, N, C, H, W, numerics, warped, valid, NULL); /* :299-305 */
    orc_homography_sample_inverse(Hts, N, H, W, flow_planes);                    /* :307-313 */

#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; n++) {
        const float *w = warped + (size_t)n * C * HW;
        memcpy(t_rgb + (size_t)n * 3 * HW, w, sizeof(float) * 3 * HW); /* :317-326 */
        memcpy(t_xyz + (size_t)n * 3 * HW, w + 4 * HW, sizeof(float) * 3 * HW);
        memcpy(t_obj + (size_t)n * HW, w + 7 * HW, sizeof(float) * HW);
        for (size_t p = 0; p < HW; p++) {
            const float z = w[6 * HW + p];
            t_sig[(size_t)n * HW + p] = (z >= 0.f) ? w[3 * HW + p] : 0.f; /* :334-338 */
            flow_chw[((size_t)n * 2 + 0) * HW + p] = flow_planes[((size_t)n * HW + p) * 2 + 0]; /* :315 */
            flow_chw[((size_t)n * 2 + 1) * HW + p] = flow_planes[((size_t)n * HW + p) * 2 + 1];
        }
    }
    /* render(): :339-345 -> :7-39.  The target-frame flow/obj composite uses the per-plane flow
     * as "src_flow" (its flow result is discarded by the caller), then the source-frame flow. */
    orc_plane_volume_rendering(t_rgb, t_sig, t_xyz, flow_chw, t_obj, B, S, H, W, is_bg_depth_inf,
                               out_rgb, out_depth, NULL, NULL, NULL, out_obj);
    orc_plane_volume_rendering_flow(mpi_sigma, flow_chw, xyz_src, B, S, H, W, hard_flow, out_flow);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++)
        for (size_t p = 0; p < HW; p++) { /* :347 */
            float cnt = 0.f;
            for (int s = 0; s < S; s++) cnt += valid[((size_t)b * S + s) * HW + p] ? 1.f : 0.f;
            out_mask[(size_t)b * HW + p] = cnt;
        }
done:
    free(depth); free(Gn); free(Kin); free(Ktn); free(Hts); free(Hst); free(cat); free(warped);
    free(valid); free(flow_planes); free(flow_chw); free(t_rgb); free(t_sig); free(t_xyz); free(t_obj);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* a6: depth-aware forward splat  (external/forward_warping/warping.c:6-33)                    */
/* ------------------------------------------------------------------------------------------ */

/* src u8[h,w,3], idx/idy i64[h,w], z f32[h,w], warped u8[h,w,5] (pre-zeroed by the caller;
 * channels: r,g,b,valid,collision).  Raster-order, last-writer z-LUT semantics. */
void orc_forward_warping(const uint8_t *src, const int64_t *idx, const int64_t *idy, const float *z,
                         uint8_t *warped, int h, int w)
{
    const size_t n = (size_t)h * w;
    float *zlut = malloc(sizeof(float) * n);
    for (size_t k = 0; k < n; k++) zlut[k] = 1000.f; /* :8-11 */
    for (size_t k = 0; k < n; k++) {                   /* :13-29, raster order over the SOURCE */
        const int x = (int)idx[k], y = (int)idy[k];    /* :16-17, long narrowed to int */
        const size_t t = (size_t)y * w + x;
        uint8_t *o = warped + t * 5;
        if (z[k] < zlut[t]) { /* :19 */
            o[0] = src[k * 3 + 0];
            o[1] = src[k * 3 + 1];
            o[2] = src[k * 3 + 2];
        }
        o[3] = 1;                              /* :23 */
        o[4] = (zlut[t] != 1000.f) ? 0 : 1;    /* :24-27 */
        zlut[t] = z[k];                        /* :28, unconditional */
    }
    free(zlut);
}

/* ------------------------------------------------------------------------------------------ */
/* a11/a12: camera helpers (geometry.py)                                                        */
/* ------------------------------------------------------------------------------------------ */

/* geometry.py:114-153 + :98-111 + :79-95.  axisangle [B,3], translation [B,3] -> M [B,4,4]. */
void orc_transformation_from_parameters(const float *axisangle, const float *translation, int B,
                                        int invert, float *M)
{
    for (int b = 0; b < B; b++) {
        const float *v = axisangle + b * 3;
        const float angle = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); /* torch.norm :119 */
        const float x = v[0] / (angle + 1e-7f), y = v[1] / (angle + 1e-7f), z = v[2] / (angle + 1e-7f);
        const float ca = cosf(angle), sa = sinf(angle), Cc = 1.f - ca;
        const float xs = x * sa, ys = y * sa, zs = z * sa, xC = x * Cc, yC = y * Cc, zC = z * Cc;
        const float xyC = x * yC, yzC = y * zC, zxC = z * xC;
        float R[16] = {x * xC + ca, xyC - zs, zxC + ys, 0.f,
                       xyC + zs, y * yC + ca, yzC - xs, 0.f,
                       zxC - ys, yzC + xs, z * zC + ca, 0.f,
                       0.f, 0.f, 0.f, 1.f};
        float t[3] = {translation[b * 3], translation[b * 3 + 1], translation[b * 3 + 2]};
        if (invert) { /* :85-87 */
            float Rt[16];
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++) Rt[i * 4 + j] = R[j * 4 + i];
            memcpy(R, Rt, sizeof(R));
            for (int i = 0; i < 3; i++) t[i] *= -1.f;
        }
        float T[16] = {1, 0, 0, t[0], 0, 1, 0, t[1], 0, 0, 1, t[2], 0, 0, 0, 1};
        const float *A = invert ? R : T, *Bm = invert ? T : R; /* :91-94 */
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) M[b * 16 + i * 4 + j] = dotk(A + i * 4, 1, Bm + j, 4, 4, 0);
    }
}

int orc_version(void) { return 1; }
