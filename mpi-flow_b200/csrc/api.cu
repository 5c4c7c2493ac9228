// api.cu -- library-level entry points: version, error strings and the host-buffer renderer.
#include <mutex>

#include "common.cuh"

MPIFLOW_API int mpiflow_version(void) { return MPIFLOW_VERSION; }

MPIFLOW_API const char *mpiflow_error_string(int code)
{
    switch (code) {
        case 0: return "success";
        case MPIFLOW_E_BADARG: return "mpiflow: bad argument (null pointer, non-positive size or unsupported value)";
        case MPIFLOW_E_WORKSPACE: return "mpiflow: workspace too small";
        case MPIFLOW_E_SINGULAR: return "Matrix inverse contains nan!";
        case MPIFLOW_E_NODEVICE: return "mpiflow: no CUDA device";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "mpiflow: unknown error";
    }
}

namespace {
// State of the host-buffer entry point: one staging buffer and two copy streams PER DEVICE (the device that is current
// when the call is made), created on first use, serialised by a per-device mutex, released by mpiflow_shutdown().
constexpr int MF_MAX_DEVICES = 64;
struct HostCtx {
    std::mutex mu;
    void *buf = nullptr;
    size_t cap = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
};
HostCtx g_ctx[MF_MAX_DEVICES];
inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace

void mf_splat_shutdown();   // splat.cu: staging buffers of the host-pointer splat

// Releases everything the host-pointer entry points (mpiflow_render_fused_host, mpiflow_forward_warp_host,
// forward_warping) keep between calls.  The device-pointer entry points keep nothing.
MPIFLOW_API int mpiflow_shutdown(void)
{
    int prev = 0;
    const bool have_dev = cudaGetDevice(&prev) == cudaSuccess;
    for (int d = 0; d < MF_MAX_DEVICES; d++) {
        HostCtx &c = g_ctx[d];
        std::lock_guard<std::mutex> lock(c.mu);
        if (!c.buf && !c.streams[0] && !c.streams[1]) continue;
        cudaSetDevice(d);
        for (int i = 0; i < 2; i++)
            if (c.streams[i]) { cudaStreamSynchronize(c.streams[i]); cudaStreamDestroy(c.streams[i]); c.streams[i] = nullptr; }
        if (c.buf) cudaFree(c.buf);
        c.buf = nullptr; c.cap = 0;
    }
    mf_splat_shutdown();
    if (have_dev) cudaSetDevice(prev);
    return 0;
}

// Host-buffer variant of mpiflow_render_fused_f32: frames are double-buffered over two streams so
// the H2D of frame b+1 overlaps the kernels and D2H of frame b (pinned host memory gets full PCIe
// bandwidth; pageable memory works but is staged by the driver).
static int render_fused_host(const void *mpi_v, const void *obj_v, size_t es, const float *disparity,
                             const float *G, const float *Kinv, const float *Kt, int B, int S,
                             int H, int W, unsigned flags, float *out_rgb, float *out_depth,
                             float *out_mask, float *out_flow, float *out_obj)
{
    // es = bytes per MPI / mask element: 4 (fp32) or 2 (fp16 storage)
    const char *mpi = reinterpret_cast<const char *>(mpi_v), *obj = reinterpret_cast<const char *>(obj_v);
    MF_CHECK_ARG(mpi_v && disparity && G && Kinv && Kt && out_rgb && out_depth && out_mask && out_flow);
    MF_CHECK_ARG(B > 0 && S > 0 && H > 0 && W > 0);
    // the host path streams one dense MPI per frame: MPIFLOW_POSES_PER_MPI belongs to the device entry point
    MF_CHECK_ARG((flags & ~(MPIFLOW_BG_DEPTH_INF | MPIFLOW_HARD_FLOW)) == 0);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e == cudaErrorNoDevice ? MPIFLOW_E_NODEVICE : (int)e;
    MF_CHECK_ARG(dev >= 0 && dev < MF_MAX_DEVICES);
    HostCtx &ctx = g_ctx[dev];
    std::lock_guard<std::mutex> lock(ctx.mu);
    for (int i = 0; i < 2; i++)
        if (!ctx.streams[i] && (e = cudaStreamCreateWithFlags(&ctx.streams[i], cudaStreamNonBlocking)) != cudaSuccess)
            return e == cudaErrorNoDevice ? MPIFLOW_E_NODEVICE : (int)e;

    const size_t HW = (size_t)H * W;
    const size_t b_mpi = al256(HW * 4 * S * es), b_obj = al256(HW * es);
    const size_t b_small = al256((S + 16 + 9 + 9) * sizeof(float)), b_out = al256(HW * 8 * sizeof(float));
    const size_t b_ws = al256(mpiflow_render_workspace_bytes(1, S));
    const size_t slot = b_mpi + b_obj + b_small + b_out + b_ws;
    const size_t need = 2 * slot + al256(2 * sizeof(int));
    if (need > ctx.cap) {
        for (int i = 0; i < 2; i++) cudaStreamSynchronize(ctx.streams[i]);
        if (ctx.buf) cudaFree(ctx.buf);
        ctx.buf = nullptr; ctx.cap = 0;
        if ((e = cudaMalloc(&ctx.buf, need)) != cudaSuccess) return (int)e;
        ctx.cap = need;
    }
    char *base = reinterpret_cast<char *>(ctx.buf);
    int *d_status = reinterpret_cast<int *>(base + 2 * slot);   // one flag per stream, sticky over the frames
    // an early return must not leave copies in flight on the caller's buffers
    auto fail = [&](int rc) {
        for (int i = 0; i < 2; i++) cudaStreamSynchronize(ctx.streams[i]);
        return rc;
    };
    if ((e = cudaMemsetAsync(d_status, 0, 2 * sizeof(int), ctx.streams[0])) != cudaSuccess) return (int)e;
    if ((e = cudaStreamSynchronize(ctx.streams[0])) != cudaSuccess) return (int)e;

    for (int b = 0; b < B; b++) {
        cudaStream_t st = ctx.streams[b & 1];
        char *d = base + (size_t)(b & 1) * slot;
        char *d_mpi = d, *d_obj = d + b_mpi;
        float *d_small = (float *)(d + b_mpi + b_obj);
        float *d_disp = d_small, *d_G = d_small + S, *d_Kinv = d_G + 16, *d_Kt = d_Kinv + 9;
        float *d_out = (float *)(d + b_mpi + b_obj + b_small);
        void *d_ws = d + b_mpi + b_obj + b_small + b_out;
        if ((e = cudaMemcpyAsync(d_mpi, mpi + (size_t)b * S * 4 * HW * es, HW * 4 * S * es,
                                 cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail((int)e);
        if (obj) cudaMemcpyAsync(d_obj, obj + (size_t)b * HW * es, HW * es, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_disp, disparity + (size_t)b * S, S * sizeof(float), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_G, G + (size_t)b * 16, 16 * sizeof(float), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_Kinv, Kinv + (size_t)b * 9, 9 * sizeof(float), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_Kt, Kt + (size_t)b * 9, 9 * sizeof(float), cudaMemcpyHostToDevice, st);
        float *o_rgb = d_out, *o_depth = d_out + 3 * HW, *o_mask = d_out + 4 * HW, *o_flow = d_out + 5 * HW,
              *o_obj = d_out + 7 * HW;
        // status = nullptr: the prologue kernel only ever raises the flag it is given, so the per-stream flag set
        // once above collects every frame of that stream
        const int rc = mf_render_fused_impl(
            (const float *)d_mpi, (int64_t)S * 4 * HW, (int64_t)4 * HW, (int64_t)HW, (const float *)(d_mpi + 3 * HW * es),
            (int64_t)S * 4 * HW, (int64_t)4 * HW, obj ? (const float *)d_obj : nullptr, (int64_t)HW, 0, d_disp, d_G, d_Kinv,
            d_Kt, 1, S, H, W, flags, o_rgb, o_depth, o_mask, o_flow, (obj && out_obj) ? o_obj : nullptr, d_ws, b_ws,
            d_status + (b & 1), /*clear_status=*/false, /*f16=*/es == 2, st);
        if (rc) return fail(rc);
        cudaMemcpyAsync(out_rgb + (size_t)b * 3 * HW, o_rgb, 3 * HW * sizeof(float), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_depth + (size_t)b * HW, o_depth, HW * sizeof(float), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_mask + (size_t)b * HW, o_mask, HW * sizeof(float), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(out_flow + (size_t)b * 2 * HW, o_flow, 2 * HW * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (obj && out_obj)
            cudaMemcpyAsync(out_obj + (size_t)b * HW, o_obj, HW * sizeof(float), cudaMemcpyDeviceToHost, st);
    }
    for (int i = 0; i < 2; i++)
        if ((e = cudaStreamSynchronize(ctx.streams[i])) != cudaSuccess) return fail((int)e);
    // error behaviour of homography_sampler.inverse (:24-26): raise when an inverse is NaN -- one readback
    int bad[2] = {0, 0};
    if ((e = cudaMemcpy(bad, d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) return (int)e;
    return (bad[0] | bad[1]) ? MPIFLOW_E_SINGULAR : 0;
}

MPIFLOW_API int mpiflow_render_fused_host(const float *mpi, const float *obj, const float *disparity,
                                          const float *G, const float *Kinv, const float *Kt, int B, int S,
                                          int H, int W, unsigned flags, float *out_rgb, float *out_depth,
                                          float *out_mask, float *out_flow, float *out_obj)
{
    return render_fused_host(mpi, obj, sizeof(float), disparity, G, Kinv, Kt, B, S, H, W, flags, out_rgb, out_depth,
                             out_mask, out_flow, out_obj);
}

// fp16 storage from host buffers: mpi [B,S,4,H,W] and obj [B,1,H,W] are IEEE halves -- half the bytes over PCIe, which
// is what bounds this entry point; everything else (and every output) is fp32
MPIFLOW_API int mpiflow_render_fused_host_f16(const void *mpi, const void *obj, const float *disparity,
                                              const float *G, const float *Kinv, const float *Kt, int B, int S,
                                              int H, int W, unsigned flags, float *out_rgb, float *out_depth,
                                              float *out_mask, float *out_flow, float *out_obj)
{
    return render_fused_host(mpi, obj, 2, disparity, G, Kinv, Kt, B, S, H, W, flags, out_rgb, out_depth, out_mask,
                             out_flow, out_obj);
}
