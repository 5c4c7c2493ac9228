// composite.cu -- un-fused operators of utils/mpi/mpi_rendering.py for arbitrary tensors:
// plane xyz (a2/a3), plane volume rendering (a7/a8) and alpha composition (a9).
// One thread per pixel, front-to-back scan over the plane axis held in registers; each input
// element is read once and each output written once (HBM-bound streaming kernels).
#include <type_traits>

#include "common.cuh"

// a2: get_src_xyz_from_plane_disparity (mpi_rendering.py:213-239)
__global__ void __launch_bounds__(256)
src_xyz_kernel(const float *__restrict__ Kinv, const float *__restrict__ disparity, int S, int H, int W,
               float *__restrict__ xyz)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, bs = blockIdx.z;
    if (x >= W) return;
    const float *k = Kinv + (size_t)(bs / S) * 9;
    const float depth = __frcp_rn(__ldg(disparity + bs));  // :224
    const size_t HW = (size_t)H * W, q = (size_t)y * W + x;
    float *o = xyz + (size_t)bs * 3 * HW + q;
#pragma unroll
    for (int i = 0; i < 3; i++)  // matmul [3x3]x[3xHW] (:234) then * depth (:235-236)
        o[i * HW] = __fmul_rn(mf_dot3(__ldg(k + i * 3), __ldg(k + i * 3 + 1), __ldg(k + i * 3 + 2),
                                      (float)x, (float)y), depth);
}

// a3: get_tgt_xyz_from_plane_disparity (:242-256) = transform_G_xyz (rendering_utils.py:4-23)
__global__ void __launch_bounds__(256)
tgt_xyz_kernel(const float *__restrict__ xyz_src, const float *__restrict__ G, int S, size_t HW,
               float *__restrict__ xyz_tgt)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int bs = blockIdx.y;
    if (q >= HW) return;
    const float *g = G + (size_t)(bs / S) * 16;
    const float *in = xyz_src + (size_t)bs * 3 * HW + q;
    const float a = __ldg(in), b = __ldg(in + HW), c = __ldg(in + 2 * HW);
    float *o = xyz_tgt + (size_t)bs * 3 * HW + q;
#pragma unroll
    for (int i = 0; i < 3; i++)  // k = 4 FMA chain, last term g[i][3] * 1
        o[i * HW] = __fadd_rn(__fmaf_rn(__ldg(g + i * 4 + 2), c,
                                        __fmaf_rn(__ldg(g + i * 4 + 1), b, __fmul_rn(__ldg(g + i * 4), a))),
                              __ldg(g + i * 4 + 3));
}

__device__ __forceinline__ float plane_dist(const float *xyz_bs, size_t HW, size_t q, int s, int S)
{
    if (s + 1 >= S) return 1e3f;  // :73-78
    const float *a = xyz_bs + q, *b = a + 3 * HW;  // plane s and s+1
    const float dx = __ldg(b) - __ldg(a), dy = __ldg(b + HW) - __ldg(a + HW),
                dz = __ldg(b + 2 * HW) - __ldg(a + 2 * HW);
    // torch.norm(dim=2) (:68-71) on CUDA: ((x^2 + y^2) + z^2), every square rounded (tools/probe_norm.py), sqrt.rn
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// planes 0..S-1 in order, plane s feeding partial sum s % 4 (the remainder of S / 4 partial sum 0)
template <class F>
__device__ __forceinline__ void planes_in_turns(int S, F &plane)
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

// a7: plane_volume_rendering + weighted_sum_mpi (:62-99, :142-154)
__global__ void __launch_bounds__(256)
plane_volume_rendering_kernel(const float *__restrict__ rgb, const float *__restrict__ sigma,
                              const float *__restrict__ xyz, const float *__restrict__ src_flow,
                              const float *__restrict__ obj, int S, size_t HW, unsigned flags,
                              float *__restrict__ rgb_out, float *__restrict__ depth_out,
                              float *__restrict__ tacc_out, float *__restrict__ w_out,
                              float *__restrict__ flow_out, float *__restrict__ obj_out)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (q >= HW) return;
    // flow: four partial sums (planes s % 4), combined at the end -- the rounding noise of one sequential fp32 chain over
    // 64+ planes exceeds 1e-4 at |flow| ~ 200, that of the reference's blocked torch.sum does not (render_pk.cuh)
    float acc = 1.f, wsum = 0.f, r = 0.f, g = 0.f, bl = 0.f, zsum = 0.f, om = 0.f;
    float fu[4] = {0.f, 0.f, 0.f, 0.f}, fv[4] = {0.f, 0.f, 0.f, 0.f};
    auto plane = [&](int s, auto acc_tag) {
        constexpr int A = decltype(acc_tag)::value;
        const size_t bs = (size_t)b * S + s;
        const float dist = plane_dist(xyz + bs * 3 * HW, HW, q, s, S);
        const float T = expf(-__ldg(sigma + bs * HW + q) * dist);  // :79
        const float w = acc * (1.f - T);                           // :80, :89
        if (tacc_out) tacc_out[bs * HW + q] = acc;                  // :84-87
        if (w_out) w_out[bs * HW + q] = w;
        acc *= (T + 1e-6f);
        wsum += w;
        const float *c = rgb + bs * 3 * HW + q;
        r = fmaf(w, __ldg(c), r);
        g = fmaf(w, __ldg(c + HW), g);
        bl = fmaf(w, __ldg(c + 2 * HW), bl);
        zsum = fmaf(w, __ldg(xyz + bs * 3 * HW + 2 * HW + q), zsum);
        if (src_flow) {  // :92-96
            fu[A] = fmaf(w, __ldg(src_flow + bs * 2 * HW + q), fu[A]);
            fv[A] = fmaf(w, __ldg(src_flow + bs * 2 * HW + HW + q), fv[A]);
            om = fmaf(w, __ldg(obj + bs * HW + q), om);
        }
    };
    planes_in_turns(S, plane);
    if (rgb_out) {
        float *o = rgb_out + (size_t)b * 3 * HW + q;
        o[0] = r; o[HW] = g; o[2 * HW] = bl;
    }
    if (depth_out)
        depth_out[(size_t)b * HW + q] = (flags & MPIFLOW_BG_DEPTH_INF) ? fmaf(1.f - wsum, 1000.f, zsum)
                                                                       : __fdiv_rn(zsum, wsum + 1e-5f);
    if (src_flow && flow_out) {
        flow_out[(size_t)b * 2 * HW + q] = __fadd_rn(__fadd_rn(fu[0], fu[1]), __fadd_rn(fu[2], fu[3]));
        flow_out[(size_t)b * 2 * HW + HW + q] = __fadd_rn(__fadd_rn(fv[0], fv[1]), __fadd_rn(fv[2], fv[3]));
    }
    if (src_flow && obj_out) obj_out[(size_t)b * HW + q] = om;
}

// a8: plane_volume_rendering_flow (:102-139)
__global__ void __launch_bounds__(256)
plane_volume_rendering_flow_kernel(const float *__restrict__ sigma, const float *__restrict__ flow,
                                   const float *__restrict__ xyz, int S, size_t HW, unsigned flags,
                                   float *__restrict__ flow_out)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (q >= HW) return;
    const bool hard = (flags & MPIFLOW_HARD_FLOW) != 0;
    float acc = 1.f, best = -1.f, fu[4] = {0.f, 0.f, 0.f, 0.f}, fv[4] = {0.f, 0.f, 0.f, 0.f};
    auto plane = [&](int s, auto acc_tag) {
        constexpr int A = decltype(acc_tag)::value;
        const size_t bs = (size_t)b * S + s;
        const float dist = plane_dist(xyz + bs * 3 * HW, HW, q, s, S);
        const float T = expf(-__ldg(sigma + bs * HW + q) * dist);
        const float w = acc * (1.f - T);
        acc *= (T + 1e-6f);
        const float u = __ldg(flow + bs * 2 * HW + q), v = __ldg(flow + bs * 2 * HW + HW + q);
        if (hard) {  // first arg-max (:126-130)
            if (w > best) { best = w; fu[0] = u; fv[0] = v; }
        } else {
            fu[A] = fmaf(w, u, fu[A]);
            fv[A] = fmaf(w, v, fv[A]);
        }
    };
    planes_in_turns(S, plane);
    flow_out[(size_t)b * 2 * HW + q] = hard ? fu[0] : __fadd_rn(__fadd_rn(fu[0], fu[1]), __fadd_rn(fu[2], fu[3]));
    flow_out[(size_t)b * 2 * HW + HW + q] = hard ? fv[0] : __fadd_rn(__fadd_rn(fv[0], fv[1]), __fadd_rn(fv[2], fv[3]));
}

// a9: alpha_composition (:42-59), C <= 8 channels per launch (host loops over channel groups)
template <int C>
__global__ void __launch_bounds__(256)
alpha_composition_kernel(const float *__restrict__ alpha, const float *__restrict__ value, int K, int Ctot,
                         int c0, size_t HW, float *__restrict__ composed, float *__restrict__ weights)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (q >= HW) return;
    float preserve = 1.f, accum[C];
#pragma unroll
    for (int c = 0; c < C; c++) accum[c] = 0.f;
    for (int k = 0; k < K; k++) {
        const size_t bk = (size_t)b * K + k;
        const float a = __ldg(alpha + bk * HW + q);
        const float w = a * preserve;  // :56
        if (weights && c0 == 0) weights[bk * HW + q] = w;
#pragma unroll
        for (int c = 0; c < C; c++)
            if (c0 + c < Ctot) accum[c] = fmaf(__ldg(value + (bk * Ctot + c0 + c) * HW + q), w, accum[c]);
        preserve *= (1.f - a);  // cumprod(1 - alpha) (:52)
    }
#pragma unroll
    for (int c = 0; c < C; c++)
        if (c0 + c < Ctot) composed[((size_t)b * Ctot + c0 + c) * HW + q] = accum[c];
}

// ---------------------------------------------------------------------------------------------
MPIFLOW_API int mpiflow_src_xyz_f32(const float *Kinv, const float *disparity, int B, int S, int H, int W,
                                    float *xyz, void *stream)
{
    MF_CHECK_ARG(Kinv && disparity && xyz && B > 0 && S > 0 && H > 0 && W > 0);
    MF_CHECK_ARG((int64_t)B * S <= 65535 && H <= 65535);
    dim3 grid((W + 255) / 256, H, B * S);
    src_xyz_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Kinv, disparity, S, H, W, xyz);
    return mf_launch_status();
}

MPIFLOW_API int mpiflow_tgt_xyz_f32(const float *xyz_src, const float *G, int B, int S, int H, int W,
                                    float *xyz_tgt, void *stream)
{
    MF_CHECK_ARG(xyz_src && G && xyz_tgt && B > 0 && S > 0 && H > 0 && W > 0 && (int64_t)B * S <= 65535);
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B * S);
    tgt_xyz_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz_src, G, S, HW, xyz_tgt);
    return mf_launch_status();
}

MPIFLOW_API int mpiflow_plane_volume_rendering_f32(const float *rgb, const float *sigma, const float *xyz,
                                                   const float *src_flow, const float *obj, int B, int S,
                                                   int H, int W, unsigned flags, float *rgb_out,
                                                   float *depth_out, float *transparency_acc,
                                                   float *weights, float *flow_out, float *obj_out,
                                                   void *stream)
{
    MF_CHECK_ARG(rgb && sigma && xyz && B > 0 && S > 0 && H > 0 && W > 0 && B <= 65535);
    MF_CHECK_ARG((src_flow == nullptr) == (obj == nullptr));
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B);
    plane_volume_rendering_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        rgb, sigma, xyz, src_flow, obj, S, HW, flags, rgb_out, depth_out, transparency_acc, weights,
        flow_out, obj_out);
    return mf_launch_status();
}

MPIFLOW_API int mpiflow_plane_volume_rendering_flow_f32(const float *src_sigma, const float *src_flow,
                                                        const float *src_xyz, int B, int S, int H, int W,
                                                        unsigned flags, float *flow_out, void *stream)
{
    MF_CHECK_ARG(src_sigma && src_flow && src_xyz && flow_out && B > 0 && S > 0 && H > 0 && W > 0 && B <= 65535);
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B);
    plane_volume_rendering_flow_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_sigma, src_flow, src_xyz,
                                                                               S, HW, flags, flow_out);
    return mf_launch_status();
}

MPIFLOW_API int mpiflow_alpha_composition_f32(const float *alpha, const float *value, int B, int K, int C,
                                              int H, int W, float *composed, float *weights, void *stream)
{
    MF_CHECK_ARG(alpha && value && composed && B > 0 && K > 0 && C > 0 && H > 0 && W > 0 && B <= 65535);
    const size_t HW = (size_t)H * W;
    dim3 grid((unsigned)((HW + 255) / 256), B);
    for (int c0 = 0; c0 < C; c0 += 4) {
        alpha_composition_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(alpha, value, K, C, c0, HW,
                                                                            composed, weights);
        int rc = mf_launch_status();
        if (rc) return rc;
    }
    return 0;
}
