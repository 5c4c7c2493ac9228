// assemble.cu -- the two passes either side of the fused renderer in one dataset sample
// (utils/utils.py:159-288, render_3dphoto_dynamic):
//   blend_src_kernel     : source-view transmittance of the MPI and the blend of the input photograph into the
//                          plane colours (:189-204) in one pass: reads 4 channels, writes 3, no xyz / weight tensors
//   assemble_sample_kernel: flow clip (:347), thresholds and merges of the object-pose and camera-pose renders
//                          (:253-283) -> flow_mix f32 [H,W,2], frame_mix u8 BGR [H,W,3], fill_mask u8 [H,W]
// Both are streaming kernels, one thread per pixel; HBM-bound.
#include "common.cuh"
#include <cstdlib>
#include <cuda_fp16.h>

// storage types of the MPI: fp32, or fp16 converted to fp32 on load (the arithmetic is fp32 either way)
__device__ __forceinline__ float mf_load(const float *p) { return __ldg(p); }
__device__ __forceinline__ float mf_load(const __half *p) { return __half2float(__ldg(p)); }

// blend_weights of mpi_rendering.render in the source frame (plane_volume_rendering, mpi_rendering.py:62-99 with
// xyz = K^-1 (x,y,1) / disparity, :213-239): the exclusive running product of (T + 1e-6), T = exp(-sigma * dist).
// Operation order as in the un-fused operators of composite.cu (src_xyz_kernel, plane_dist), which the torch
// restatement is checked against; then  w * src + (1 - w) * rgb  as four separately rounded operations (:195-197).
// OUT_C = 3: out [B,S,3,H,W]; OUT_C = 4: out [B,S,4,H,W] with sigma (converted to fp32) as the fourth channel, i.e. the
// MPI the renderer takes, so that an fp16 MPI is read exactly once on its way to fp32.
template <int UNROLL, class TIn, int OUT_C>
__global__ void __launch_bounds__(256)
blend_src_kernel(const TIn *__restrict__ rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc,
                 const TIn *__restrict__ sigma, int64_t sig_sb, int64_t sig_ss,
                 const TIn *__restrict__ src, const float *__restrict__ disparity, const float *__restrict__ Kinv,
                 int S, int H, int W, float *__restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const size_t HW = (size_t)H * W, q = (size_t)y * W + x;
    const float *k = Kinv + (size_t)b * 9;
    float ray[3], px[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        ray[i] = mf_dot3(__ldg(k + i * 3), __ldg(k + i * 3 + 1), __ldg(k + i * 3 + 2), (float)x, (float)y);
        px[i] = mf_load(src + ((size_t)b * 3 + i) * HW + q);
    }
    const float *disp = disparity + (size_t)b * S;
    const TIn *rg = rgb + (size_t)b * rgb_sb + q, *sg = sigma + (size_t)b * sig_sb + q;
    float *o = out + (size_t)b * S * OUT_C * HW + q;
    float acc = 1.f;
    float dcur = __frcp_rn(__ldg(disp));
    for (int s0 = 0; s0 < S; s0 += UNROLL) {
        float sv[UNROLL], cv[UNROLL][3];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {   // all loads of the group first: UNROLL * 4 requests in flight per thread
            const int s = min(s0 + u, S - 1);
            sv[u] = mf_load(sg + (size_t)s * sig_ss);
#pragma unroll
            for (int c = 0; c < 3; c++) cv[u][c] = mf_load(rg + (size_t)s * rgb_ss + (size_t)c * rgb_sc);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int s = s0 + u;
            if (s >= S) break;
            float dist = 1e3f;
            if (s + 1 < S) {
                const float dnext = __frcp_rn(__ldg(disp + s + 1));
                const float ex = __fsub_rn(__fmul_rn(ray[0], dnext), __fmul_rn(ray[0], dcur));
                const float ey = __fsub_rn(__fmul_rn(ray[1], dnext), __fmul_rn(ray[1], dcur));
                const float ez = __fsub_rn(__fmul_rn(ray[2], dnext), __fmul_rn(ray[2], dcur));
                // torch.norm on CUDA: ((x^2 + y^2) + z^2) with every square rounded (tools/probe_norm.py)
                dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
                dcur = dnext;
            }
            const float T = expf(__fmul_rn(-sv[u], dist));
            const float keep = __fsub_rn(1.f, acc);
#pragma unroll
            for (int c = 0; c < 3; c++)
                o[((size_t)s * OUT_C + c) * HW] = __fadd_rn(__fmul_rn(acc, px[c]), __fmul_rn(keep, cv[u][c]));
            if (OUT_C == 4) o[((size_t)s * OUT_C + 3) * HW] = sv[u];
            acc = __fmul_rn(acc, __fadd_rn(T, 1e-6f));
        }
    }
}

template <class TIn, int OUT_C>
static int launch_blend(const TIn *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc, const TIn *sigma, int64_t sig_sb,
                        int64_t sig_ss, const TIn *src_img, const float *disparity, const float *Kinv, int B, int S, int H,
                        int W, float *out, void *stream)
{
    if (!rgb || !sigma || !src_img || !disparity || !Kinv || !out) return MPIFLOW_E_BADARG;
    if (B <= 0 || S <= 0 || H <= 0 || W <= 0 || H > 65535 || B > 65535) return MPIFLOW_E_BADARG;
    static const int unroll = []() {   // MPIFLOW_BLEND_UNROLL=2|4|8: planes loaded ahead per thread (A/B measurements)
        const char *e = getenv("MPIFLOW_BLEND_UNROLL");
        return e ? atoi(e) : 4;
    }();
    static const int block = []() {
        const char *e = getenv("MPIFLOW_BLEND_BLOCK");
        const int v = e ? atoi(e) : 256;
        return (v == 64 || v == 128 || v == 256) ? v : 256;
    }();
    dim3 grid((W + block - 1) / block, H, B);
    auto kern = unroll == 8 ? blend_src_kernel<8, TIn, OUT_C> : unroll == 2 ? blend_src_kernel<2, TIn, OUT_C>
                                                                             : blend_src_kernel<4, TIn, OUT_C>;
    kern<<<grid, block, 0, (cudaStream_t)stream>>>(rgb, rgb_sb, rgb_ss, rgb_sc, sigma, sig_sb, sig_ss, src_img, disparity,
                                                   Kinv, S, H, W, out);
    return mf_launch_status();
}

MPIFLOW_API int mpiflow_blend_src_f32(const float *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc,
                                      const float *sigma, int64_t sig_sb, int64_t sig_ss, const float *src_img,
                                      const float *disparity, const float *Kinv, int B, int S, int H, int W,
                                      float *out_rgb, void *stream)
{
    return launch_blend<float, 3>(rgb, rgb_sb, rgb_ss, rgb_sc, sigma, sig_sb, sig_ss, src_img, disparity, Kinv, B, S, H, W,
                                  out_rgb, stream);
}

// fp16 storage (the generator keeps its MPI and photograph in half precision, gen_3dphoto_dynamic_v2.py:82-93): the same
// blend with fp32 arithmetic on the fp16 values, written as the fp32 MPI [B,S,4,H,W] (blended colours + sigma) that
// mpiflow_render_fused_f32 takes -- the half MPI is read once and never up-cast by a separate pass.
MPIFLOW_API int mpiflow_blend_src_f16(const void *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc,
                                      const void *sigma, int64_t sig_sb, int64_t sig_ss, const void *src_img,
                                      const float *disparity, const float *Kinv, int B, int S, int H, int W,
                                      float *out_mpi, void *stream)
{
    return launch_blend<__half, 4>((const __half *)rgb, rgb_sb, rgb_ss, rgb_sc, (const __half *)sigma, sig_sb, sig_ss,
                                   (const __half *)src_img, disparity, Kinv, B, S, H, W, out_mpi, stream);
}

__device__ __forceinline__ unsigned char to_u8(float v)
{
    // np.clip(np.round(v * 255), 0, 255).astype(np.uint8): round half to even in fp32 (:241-250)
    return (unsigned char)fminf(fmaxf(rintf(__fmul_rn(v, 255.f)), 0.f), 255.f);
}

// pose 0 = the object-mask render ("frame", camera pose with base motion 0), pose 1 = the inverse-mask render
// ("frame_dync"); channel-planar fp32 in, pixel-interleaved out as the reference hands them to OpenCV / writeFlow
__global__ void __launch_bounds__(256)
assemble_sample_kernel(const float *__restrict__ frame, const float *__restrict__ frame_dync,
                       const float *__restrict__ mask, const float *__restrict__ mask_dync,
                       const float *__restrict__ flow, const float *__restrict__ flow_dync,
                       const float *__restrict__ obj, float thresh, size_t HW, float *__restrict__ flow_mix,
                       unsigned char *__restrict__ frame_mix, unsigned char *__restrict__ fill_mask)
{
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= HW) return;
    const float m = __ldg(mask + q), md = __ldg(mask_dync + q);
    const bool own = m >= thresh;                                 // :275, :281
    const bool from_obj = __ldg(obj + q) >= thresh;               // :268-269, :277
    const float *f = from_obj ? flow : flow_dync;
    flow_mix[2 * q] = fminf(fmaxf(__ldg(f + q), -200.f), 200.f);  // torch.clip(flow_syn, -200, 200), :347
    flow_mix[2 * q + 1] = fminf(fmaxf(__ldg(f + HW + q), -200.f), 200.f);
    const float *c = own ? frame : frame_dync;
    const bool hole = own ? false : (md < thresh);                // :271-272 (own implies mask >= thresh)
#pragma unroll
    for (int i = 0; i < 3; i++)                                    // [2,1,0]: RGB -> BGR
        frame_mix[3 * q + i] = hole ? (unsigned char)255 : to_u8(__ldg(c + (size_t)(2 - i) * HW + q));
    fill_mask[q] = (unsigned char)(((own ? 1.f : md) < thresh) ? 1 : 0);   // :279-283
}

MPIFLOW_API int mpiflow_assemble_sample_f32(const float *frame, const float *frame_dync, const float *mask,
                                            const float *mask_dync, const float *flow, const float *flow_dync,
                                            const float *obj_mask, float mask_thresh, int H, int W,
                                            float *flow_mix, unsigned char *frame_mix, unsigned char *fill_mask,
                                            void *stream)
{
    if (!frame || !frame_dync || !mask || !mask_dync || !flow || !flow_dync || !obj_mask || !flow_mix || !frame_mix ||
        !fill_mask)
        return MPIFLOW_E_BADARG;
    if (H <= 0 || W <= 0) return MPIFLOW_E_BADARG;
    const size_t HW = (size_t)H * W;
    assemble_sample_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        frame, frame_dync, mask, mask_dync, flow, flow_dync, obj_mask, mask_thresh, HW, flow_mix, frame_mix, fill_mask);
    return mf_launch_status();
}
