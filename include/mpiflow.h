/*
 * mpiflow.h -- C ABI of libmpiflow_sm100.so: the B200 (sm_100a) implementation of MPI-Flow's
 * multiplane-image renderer hot path and depth-aware forward splat.
 *
 * Conventions
 *   - every entry point is extern "C", takes plain pointers / sizes, returns int:
 *       0 = success, >0 = a cudaError_t, <0 = MPIFLOW_E_* argument error.  Nothing throws.
 *   - "_f32" entry points take DEVICE pointers and a cudaStream_t (passed as void*); they never
 *     allocate, never synchronise and keep no global state: scratch comes from the caller
 *     (size from the matching *_workspace_bytes()).
 *   - "_host" entry points (and the legacy forward_warping) take HOST pointers, do their own
 *     H2D/D2H and return when the result is in the caller's buffers.
 *   - tensors are fp32, the two innermost dims (H, W) contiguous; outer strides are given in
 *     ELEMENTS so that torch views (mpi[:, :, 0:3], expand()ed masks) are accepted as they are.
 *   - reference citations are relative to the MPI-Flow repository root.
 */
#ifndef MPIFLOW_H_
#define MPIFLOW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPIFLOW_VERSION 100

#define MPIFLOW_E_BADARG (-1)    /* null pointer / non-positive size / unsupported value */
#define MPIFLOW_E_WORKSPACE (-2) /* workspace too small */
#define MPIFLOW_E_SINGULAR (-3)  /* "Matrix inverse contains nan!" (homography_sampler.py:24-26) */
#define MPIFLOW_E_NODEVICE (-4)  /* no sm_100 device / CUDA not initialised */

/* render flags (keyword arguments of mpi_rendering.render_tgt_rgb_depth, mpi_rendering.py:259-272) */
#define MPIFLOW_BG_DEPTH_INF 1u /* is_bg_depth_inf */
#define MPIFLOW_HARD_FLOW 2u    /* hard_flow */
/* The caller promises that nobody else writes to `workspace` between calls (a workspace owned by one loop, as the
 * generator's): the plane-spacing table of the previous call is then reused whenever K_src^-1 of entry 0, the
 * disparities, S, H and W are bit for bit what they were (checked on the device against a key stored behind the table;
 * no host sync), and launches of fewer than 4 frames use the table too.  Needs mpiflow_render_workspace_bytes_hw.
 * Zero a workspace once before its first call with this flag (recycled memory may hold the key of an earlier table). */
#define MPIFLOW_KEEP_TABLE 4u
/* n consecutive frames of the batch render the same MPI under n poses (utils/utils.py:210-238 renders every sample
 * twice): rgb / sigma then hold B/n MPIs, rgb_sb / sig_sb step between MPIs, everything else stays per frame.
 * n = 1..256, B % n == 0.  (A batch stride of 0 says the same for "all B frames share one MPI".) */
#define MPIFLOW_POSES_PER_MPI(n) ((((unsigned)(n) - 1u) & 0xffu) << 8)
/* the persistent renderer occupies every SM for the whole launch; n > 0 leaves n SMs free for kernels of other streams
 * that must make progress meanwhile (a collective gathering the previous step's outputs, shard.FrameGatherer). n = 0..255 */
#define MPIFLOW_RESERVE_SMS(n) (((unsigned)(n) & 0xffu) << 16)

int mpiflow_version(void);
/* frees the per-device staging buffers and streams of the host-pointer entry points (they are re-created on the next
 * call); the device-pointer entry points keep no state.  Always returns 0. */
int mpiflow_shutdown(void);
/* static string for any code returned by this library */
const char *mpiflow_error_string(int code);

/* ---- a4/a5 prologue: plane-induced homographies ------------------------------------------------
 * Replaces the host-side maths of HomographySample.sample / sample_inverse
 * (utils/mpi/homography_sampler.py:96-122, :176-197): H_tgt_src = K_tgt (R - t n^T / -d) K_src^-1 in
 * fp32 (cuBLAS accumulation order), H_src_tgt = fp32(inverse_fp64(H_tgt_src)).  The reference does the
 * inverse on the CPU with a device sync; here it is a device kernel, no sync.
 *   G [N,4,4], Kinv [N,3,3], Kt [N,3,3], depth [N]  ->  Hts [N,9], Hst [N,9]
 *   status: device int, set non-zero if any inverse is singular / NaN (may be NULL). */
int mpiflow_plane_homographies_f32(const float *G, const float *Kinv, const float *Kt, const float *depth,
                                   int N, float *Hts, float *Hst, int *status, void *stream);

/* ---- a10: fused render_tgt_rgb_depth -----------------------------------------------------------
 * One launch replaces get_src_xyz_from_plane_disparity + get_tgt_xyz_from_plane_disparity +
 * HomographySample.sample + sample_inverse + render (plane_volume_rendering +
 * plane_volume_rendering_flow) + the tgt_mask reduction: mpi_rendering.py:213-256, :259-349,
 * homography_sampler.py:80-220, for the sigma mode (use_alpha=False) with plane-derived xyz.
 *
 *   rgb    [B,S,3,H,W]  strides (rgb_sb, rgb_ss, rgb_sc)      sigma [B,S,1,H,W] strides (sig_sb, sig_ss)
 *   obj    [B,S,1,H,W]  strides (obj_sb, obj_ss); obj_ss = 0 for a mask shared by all planes
 *   disparity [B,S] (dense), G [B,4,4], Kinv [B,3,3], Kt [B,3,3]
 *   out_rgb [B,3,H,W], out_depth [B,1,H,W], out_mask [B,1,H,W] (count of valid planes, :347),
 *   out_flow [B,2,H,W] (flowA2B, source frame), out_obj [B,1,H,W]      -- all dense
 *   workspace: mpiflow_render_workspace_bytes(B, S) bytes, device, 16-byte aligned; with
 *     mpiflow_render_workspace_bytes_hw(B, S, H, W) bytes (S*H*W floats and a key more) launches of 4 or more frames whose
 *     K_src^-1 and disparities agree evaluate the source-frame plane spacing once per launch instead of per frame
 *     (MPIFLOW_KEEP_TABLE: once per workspace, and for any number of frames)
 *   status: device int (may be NULL), non-zero when a homography was singular */
size_t mpiflow_render_workspace_bytes(int B, int S);
size_t mpiflow_render_workspace_bytes_hw(int B, int S, int H, int W);
int mpiflow_render_fused_f32(const float *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc,
                             const float *sigma, int64_t sig_sb, int64_t sig_ss,
                             const float *obj, int64_t obj_sb, int64_t obj_ss,
                             const float *disparity, const float *G, const float *Kinv, const float *Kt,
                             int B, int S, int H, int W, unsigned flags,
                             float *out_rgb, float *out_depth, float *out_mask, float *out_flow,
                             float *out_obj, void *workspace, size_t workspace_bytes, int *status,
                             void *stream);

/* fp16 STORAGE of the same computation: rgb / sigma / obj are IEEE halves (strides in elements), every other argument
 * and every output as in mpiflow_render_fused_f32; the arithmetic is fp32 on the converted values (parity: identical
 * to the fp32 entry on the up-cast tensors to rounding of the accumulation order, 1e-4).  Half the HBM / PCIe bytes per
 * MPI.  Served by the plain global-load kernel (no TMA staging): meant for callers bounded by moving the MPI. */
int mpiflow_render_fused_f16(const void *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc,
                             const void *sigma, int64_t sig_sb, int64_t sig_ss,
                             const void *obj, int64_t obj_sb, int64_t obj_ss,
                             const float *disparity, const float *G, const float *Kinv, const float *Kt,
                             int B, int S, int H, int W, unsigned flags,
                             float *out_rgb, float *out_depth, float *out_mask, float *out_flow,
                             float *out_obj, void *workspace, size_t workspace_bytes, int *status,
                             void *stream);

/* test hook: the shared-reciprocal IEEE division of the kernels, element-wise: qa = a/b, qc = c/b */
int mpiflow_debug_div_f32(const float *a, const float *c, const float *b, int n, float *qa, float *qc,
                          void *stream);

/* test hook: packed fp32x2 arithmetic of the persistent renderer next to the scalar IEEE operations; n even,
 * out [14,n]: rows fma2, fma, mul2, mul, sub2, sub, floor via 2^23 (round-down add), floorf, unfused a*b-1
 * packed, scalar, packed clone of libdevice expf on -|a|/22, expf of the same,
 * packed IEEE sqrt of |a*b|, sqrtf of the same */
int mpiflow_debug_packed_f32(const float *a, const float *b, const float *c, int n, float *out, void *stream);

/* Same computation from HOST buffers (dense layouts: mpi [B,S,4,H,W] = rgb+sigma, obj [B,1,H,W]):
 * H2D of the inputs, the two launches, D2H of the five outputs, all inside the call; returns
 * MPIFLOW_E_SINGULAR like the reference raises.  Frames are pipelined over two streams.
 * Unlike the device-pointer entry points, the host-pointer ones (this one, mpiflow_forward_warp_host,
 * forward_warping) KEEP STATE between calls: a device staging buffer and two streams per CUDA device (the device
 * current at the call), created on first use, guarded by a per-device mutex (concurrent callers on one device are
 * serialised), released by mpiflow_shutdown().  On an error return no copy is left in flight on the caller's buffers. */
int mpiflow_render_fused_host(const float *mpi, const float *obj, const float *disparity, const float *G,
                              const float *Kinv, const float *Kt, int B, int S, int H, int W,
                              unsigned flags, float *out_rgb, float *out_depth, float *out_mask,
                              float *out_flow, float *out_obj);

/* the same from host buffers holding IEEE halves: mpi [B,S,4,H,W], obj [B,1,H,W] (may be NULL) */
int mpiflow_render_fused_host_f16(const void *mpi, const void *obj, const float *disparity, const float *G,
                                  const float *Kinv, const float *Kt, int B, int S, int H, int W,
                                  unsigned flags, float *out_rgb, float *out_depth, float *out_mask,
                                  float *out_flow, float *out_obj);

/* ---- fine-grained operators (same maths, un-fused, arbitrary tensors) ---------------------------- */
/* a4  HomographySample.sample (homography_sampler.py:124-158): src [N,C,H,W] (dense) warped by Hst
 *     [N,9] -> tgt [N,C,H,W], valid [N,H,W] (u8 0/1), flowB2A [N,H,W,2] (either may be NULL) */
int mpiflow_homography_sample_f32(const float *src, const float *Hst, int N, int C, int H, int W,
                                  float *tgt, uint8_t *valid, float *flowB2A, void *stream);
/* a5  HomographySample.sample_inverse (:199-220): Hts [N,9] -> flowA2B [N,H,W,2] */
int mpiflow_homography_flow_f32(const float *Hts, int N, int H, int W, float *flowA2B, void *stream);
/* a2  get_src_xyz_from_plane_disparity (mpi_rendering.py:213-239) -> xyz [B,S,3,H,W] */
int mpiflow_src_xyz_f32(const float *Kinv, const float *disparity, int B, int S, int H, int W,
                        float *xyz, void *stream);
/* a3  get_tgt_xyz_from_plane_disparity (:242-256) / transform_G_xyz (rendering_utils.py:4-23) */
int mpiflow_tgt_xyz_f32(const float *xyz_src, const float *G, int B, int S, int H, int W,
                        float *xyz_tgt, void *stream);
/* a7  plane_volume_rendering + weighted_sum_mpi (:62-99, :142-154).  rgb [B,S,3,H,W], sigma
 *     [B,S,1,H,W], xyz [B,S,3,H,W] dense; src_flow [B,S,2,H,W] / obj [B,S,1,H,W] both or neither.
 *     Any output may be NULL: rgb_out [B,3,H,W], depth_out [B,1,H,W], transparency_acc [B,S,1,H,W],
 *     weights [B,S,1,H,W], flow_out [B,2,H,W], obj_out [B,1,H,W]. */
int mpiflow_plane_volume_rendering_f32(const float *rgb, const float *sigma, const float *xyz,
                                       const float *src_flow, const float *obj, int B, int S, int H, int W,
                                       unsigned flags, float *rgb_out, float *depth_out,
                                       float *transparency_acc, float *weights, float *flow_out,
                                       float *obj_out, void *stream);
/* a8  plane_volume_rendering_flow (:102-139) -> flow_out [B,2,H,W] */
int mpiflow_plane_volume_rendering_flow_f32(const float *src_sigma, const float *src_flow,
                                            const float *src_xyz, int B, int S, int H, int W,
                                            unsigned flags, float *flow_out, void *stream);
/* a9  alpha_composition (:42-59): alpha [B,K,1,H,W], value [B,K,C,H,W] -> composed [B,C,H,W],
 *     weights [B,K,1,H,W] (may be NULL) */
int mpiflow_alpha_composition_f32(const float *alpha, const float *value, int B, int K, int C, int H,
                                  int W, float *composed, float *weights, void *stream);

/* ---- a6: depth-aware forward splat ---------------------------------------------------------------
 * Legacy symbol, identical signature and semantics to external/forward_warping/warping.c:6
 * (HOST pointers; src u8[h,w,3], idx/idy int64[h,w], z f32[h,w], warped u8[h,w,5] pre-zeroed by the
 * caller; no return value).  Bit-exact with the serial C loop, computed on the GPU.
 * Also exported by external/forward_warping/libwarping.so for moving_obj.py:12-13. */
void forward_warping(const void *src, const void *idx, const void *idy, const void *z, void *warped,
                     int h, int w);
/* like forward_warping but reports errors */
int mpiflow_forward_warp_host(const void *src, const void *idx, const void *idy, const void *z,
                              void *warped, int h, int w);
/* DEVICE pointers, same buffers; warped must hold the caller's initial content (normally zeros). */
size_t mpiflow_forward_warp_workspace_bytes(int h, int w);
int mpiflow_forward_warp_f32(const uint8_t *src, const int64_t *idx, const int64_t *idy, const float *z,
                             uint8_t *warped, int h, int w, void *workspace, size_t workspace_bytes,
                             void *stream);

/* ---- mesh rasteriser behind warpback.utils.RGBDRenderer.render_mesh ----------------
 * Replaces pytorch3d.renderer.mesh.rasterize_meshes(faces_per_pixel=1, blur_radius) followed by
 * pytorch3d.ops.interpolate_face_attributes (reference warpback/utils.py:56-63).  PARITY UNPINNED: pytorch3d is an
 * un-vendored dependency (0.6.2, README.md:30) absent from this image; the kernels follow its published naive
 * rasteriser and are checked against oracle/raster_oracle.py.
 *   verts_ndc [B,V,3] (x, y in pytorch3d NDC: +X left, +Y up, shorter image side spans [-1,1]; z = depth > 0)
 *   faces [B,F,3] int64 vertex indices, attrs [B,V,C]
 *   out [B,C,H,W] interpolated attributes (0 where no face), pix_to_face [B,H,W] int64 (b*F + f, -1 = none),
 *   zbuf [B,H,W], bary [B,H,W,3]  -- each may be NULL
 *   workspace: mpiflow_rasterize_workspace_bytes(B, H, W) bytes, 8-byte aligned */
size_t mpiflow_rasterize_workspace_bytes(int B, int H, int W);
int mpiflow_rasterize_mesh_f32(const float *verts_ndc, const int64_t *faces, const float *attrs, int B, int V, int F,
                               int C, int H, int W, float blur_radius, float *out, int64_t *pix_to_face,
                               float *zbuf, float *bary, void *workspace, size_t workspace_bytes, void *stream);

/* ---- sample assembly around the renderer (utils/utils.py:159-288) -----------------------
 * mpiflow_blend_src_f32 replaces the source-view mpi_rendering.render call and the blend of the photograph into the
 * plane colours (utils/utils.py:183-204): out[b,s,c] = w * src_img[b,c] + (1 - w) * rgb[b,s,c] with
 * w = blend_weights[b,s] = prod_{j<s} (exp(-sigma_j * dist_j) + 1e-6) in the source frame.  One pass: 4 S channels
 * read, 3 S written, no xyz / weight tensors.
 *   rgb / sigma: strided as in mpiflow_render_fused_f32; src_img [B,3,H,W]; disparity [B,S]; Kinv [B,3,3]
 *   out_rgb [B,S,3,H,W] dense, must not overlap the inputs */
int mpiflow_blend_src_f32(const float *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc, const float *sigma,
                          int64_t sig_sb, int64_t sig_ss, const float *src_img, const float *disparity,
                          const float *Kinv, int B, int S, int H, int W, float *out_rgb, void *stream);
/* fp16 storage of the same step (the generator keeps the MPI and the photograph in half precision,
 * gen_3dphoto_dynamic_v2.py:82-93): rgb / sigma / src_img are IEEE half (strides in elements), the arithmetic is fp32
 * on the converted values, and the result is the fp32 MPI mpiflow_render_fused_f32 takes:
 *   out_mpi [B,S,4,H,W] dense: channels 0..2 the blended colours, channel 3 sigma converted to fp32.
 * The half MPI is read once; no separate up-cast pass. */
int mpiflow_blend_src_f16(const void *rgb, int64_t rgb_sb, int64_t rgb_ss, int64_t rgb_sc, const void *sigma,
                          int64_t sig_sb, int64_t sig_ss, const void *src_img, const float *disparity,
                          const float *Kinv, int B, int S, int H, int W, float *out_mpi, void *stream);
/* mpiflow_assemble_sample_f32 replaces the host-side thresholds and merges of the two renders of one sample
 * (utils/utils.py:241-283, flow clip :347): frame/mask/flow = outputs of the render with obj_mask, *_dync = outputs
 * of the render with 1 - obj_mask; all channel-planar fp32 [3|1|2,H,W]; obj_mask [H,W].
 *   flow_mix f32 [H,W,2]; frame_mix u8 [H,W,3] in B,G,R order (holes 255); fill_mask u8 [H,W] (1 = to inpaint) */
int mpiflow_assemble_sample_f32(const float *frame, const float *frame_dync, const float *mask,
                                const float *mask_dync, const float *flow, const float *flow_dync,
                                const float *obj_mask, float mask_thresh, int H, int W, float *flow_mix,
                                unsigned char *frame_mix, unsigned char *fill_mask, void *stream);

/* ---- hole filling of the assembled sample ---------------------------------------------
 * mpiflow_inpaint_ns_u8 replaces cv2.inpaint(frame_mix, fill_mask, 3, cv2.INPAINT_NS) (reference utils/utils.py:284-286;
 * OpenCV is an un-vendored dependency of the reference): OpenCV's fast-marching Navier-Stokes inpainting, BIT FOR BIT
 * (same arrival order, same fp32 / fp64 steps; oracle/inpaint_oracle.c restates it and is pinned against cv2.inpaint).
 * A batch of n independent images per call, one warp marching each image.
 *   img  u8 [n,rows,cols,3] (interleaved, as cv2), mask u8 [n,rows,cols] (non-zero = to fill), radius 1..5
 *   out  u8 [n,rows,cols,3]; may be img itself (in place)
 *   workspace: mpiflow_inpaint_ns_workspace_bytes(n, rows, cols) bytes, 256-byte aligned; no allocation, no sync */
size_t mpiflow_inpaint_ns_workspace_bytes(int n, int rows, int cols);
int mpiflow_inpaint_ns_u8(const uint8_t *img, const uint8_t *mask, int n, int rows, int cols, int radius,
                          uint8_t *out, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MPIFLOW_H_ */
