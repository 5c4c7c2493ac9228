// homography.cuh -- device helpers shared by the prologue, the un-fused warp and the fused renderer.
//
// Restates, for the GPU, the host-side maths of HomographySample.sample / sample_inverse
// (utils/mpi/homography_sampler.py:96-122, :176-197) and the per-pixel coordinate chain (:124-158).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// plane homographies
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mm33_fma(const float *A, const float *B, float *C)
{
    // cuBLAS accumulation order: acc = a0*b0; acc = fma(a1,b1,acc); acc = fma(a2,b2,acc)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[i * 3 + j] = __fmaf_rn(A[i * 3 + 2], B[6 + j],
                                     __fmaf_rn(A[i * 3 + 1], B[3 + j], __fmul_rn(A[i * 3], B[j])));
}

// fp64 Gauss-Jordan with partial pivoting (the reference uses LAPACK on the CPU in fp64,
// homography_sampler.py:15-17,121-122; after rounding to fp32 the results coincide).
__device__ inline bool inverse33_f64(const double *A, double *Ainv)
{
    double m[3][6];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            m[i][j] = A[i * 3 + j];
            m[i][3 + j] = (i == j) ? 1.0 : 0.0;
        }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int p = c;
        for (int r = c + 1; r < 3; r++)
            if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (m[p][c] == 0.0) return false;
        if (p != c)
            for (int j = 0; j < 6; j++) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
        const double inv = 1.0 / m[c][c];
        for (int j = 0; j < 6; j++) m[c][j] *= inv;
        for (int r = 0; r < 3; r++) {
            if (r == c) continue;
            const double f = m[r][c];
            for (int j = 0; j < 6; j++) m[r][j] -= f * m[c][j];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Ainv[i * 3 + j] = m[i][3 + j];
    return true;
}

// Computes H_tgt_src / H_src_tgt for entry n from G (4x4), Kinv, Kt (3x3) and depth.
__device__ inline bool plane_homography(const float *g, const float *kinv, const float *kt, float depth,
                                 float *hts, float *hst)
{
    float Rtnd[9], M1[9];
    const float negd = -depth;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        // R - (t n^T) / -d with n = (0,0,1): only column 2 changes (:114-116)
        Rtnd[i * 3 + 0] = g[i * 4 + 0];
        Rtnd[i * 3 + 1] = g[i * 4 + 1];
        Rtnd[i * 3 + 2] = __fsub_rn(g[i * 4 + 2], __fdiv_rn(g[i * 4 + 3], negd));
    }
    mm33_fma(Rtnd, kinv, M1);  // :117-118
    mm33_fma(kt, M1, hts);
    double h64[9], hi64[9];
#pragma unroll
    for (int i = 0; i < 9; i++) h64[i] = (double)hts[i];
    bool ok = inverse33_f64(h64, hi64);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        hst[i] = ok ? (float)hi64[i] : __int_as_float(0x7fc00000);
        ok = ok && (hst[i] == hst[i]);
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------
// a4: HomographySample.sample, per pixel (homography_sampler.py:124-158)
// ---------------------------------------------------------------------------------------------
struct SrcPos {
    float sx, sy;  // source position of the target pixel
    float ix, iy;  // grid_sample's clipped source index
    bool valid;
};

// Target pixel (x,y) -> source position, validity, clipped index.  Every operation is the one the
// reference executes on CUDA, in its order (see DESIGN.md "numerics contract").
__device__ __forceinline__ SrcPos tgt_to_src(const float *h, float x, float y, float Wf, float Hf,
                                             float rW, float rH)
{
    SrcPos p;
    const float X = mf_dot3(h[0], h[1], h[2], x, y);
    const float Y = mf_dot3(h[3], h[4], h[5], x, y);
    const float Z = mf_dot3(h[6], h[7], h[8], x, y);
    mf_div2(X, Y, Z, p.sx, p.sy);                                           // :137-138
    p.valid = (p.sx < Wf) && (p.sx > -1.f) && (p.sy < Hf) && (p.sy > -1.f);  // :143-147
    // :151-154, tensor / python scalar on CUDA = multiply by the fp32 reciprocal
    const float gx = __fadd_rn(__fmul_rn(__fadd_rn(p.sx, 0.5f), rW), -1.f);
    const float gy = __fadd_rn(__fmul_rn(__fadd_rn(p.sy, 0.5f), rH), -1.f);
    // ATen GridSampler.cuh:29 ((g+1)*size-1)/2 as nvcc compiles it (FADD, FFMA, FMUL)
    float ix = __fmul_rn(__fmaf_rn(__fadd_rn(gx, 1.f), Wf, -1.f), 0.5f);
    float iy = __fmul_rn(__fmaf_rn(__fadd_rn(gy, 1.f), Hf, -1.f), 0.5f);
    p.ix = fminf(Wf - 1.f, fmaxf(ix, 0.f));  // padding_mode='border' (GridSampler.cuh:55-56)
    p.iy = fminf(Hf - 1.f, fmaxf(iy, 0.f));
    return p;
}

