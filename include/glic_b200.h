/*
 * glic_b200.h -- C ABI of the B200-native (sm_100a) Gaussian rasterizer hot path.
 *
 * Drop-in boundary for Gaussian-LIC (APRIL-ZJU/Gaussian-LIC @ 4566e6e).  Every entry point
 * replaces one raw-pointer interface of the reference; the LibTorch symbols the SLAM code
 * actually links against (RasterizeGaussiansCUDA, RasterizeGaussiansBackwardCUDA, adamUpdate,
 * fusedssim, fusedssim_backward, distCUDA2) are thin shims over this ABI
 * (gaussian_lic_b200/csrc/torch_shim.cpp).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types, no exceptions across the ABI.
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what
 *     the reference uses for every launch).
 *   - the library never allocates device memory: the caller provides workspaces whose sizes
 *     come from the glic_*_bytes() queries (replaces the reference's four
 *     std::function<char*(size_t)> resize callbacks, rasterizer.h:29-34).
 *   - return value: GLIC_OK (0) or a negative glic_status; glic_last_error() gives text.
 *   - all floating point is IEEE fp32; indices are 32-bit, keys 64-bit.
 *
 * Reference citations are relative to /root/reference/src/.
 */
#ifndef GLIC_B200_H_INCLUDED
#define GLIC_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLIC_ABI_VERSION 1

#if defined(__GNUC__)
#define GLIC_API __attribute__((visibility("default")))
#else
#define GLIC_API
#endif

typedef enum glic_status {
    GLIC_OK = 0,
    GLIC_ERR_INVALID_ARGUMENT = -1, /* bad shape / null pointer / unsupported degree        */
    GLIC_ERR_WORKSPACE = -2,        /* a workspace is smaller than glic_*_bytes() requires   */
    GLIC_ERR_CUDA = -3,             /* a CUDA runtime call failed (text in glic_last_error)  */
    GLIC_ERR_NO_DEVICE = -4,        /* no CUDA device / wrong architecture                   */
    GLIC_ERR_TIMEOUT = -5           /* a multi-GPU peer did not reach the exchange barrier   */
} glic_status;

/* Tile geometry is part of the parity contract (rasterizer/cuda_rasterizer/config.h:16-17). */
#define GLIC_TILE 16
#define GLIC_BUCKET 32 /* checkpoint period of the per-splat backward (forward.cu:412) */

GLIC_API const char* glic_last_error(void);
GLIC_API int glic_abi_version(void);

/* Per-view camera block, all DEVICE pointers except the scalars.
 * view/proj: 16 floats each, column-major Rt and P*Rt exactly as the reference's kernels get
 * them (rasterize_points.cu:129-130, camera.h:60,86,109).  campos: 3 floats. */
typedef struct glic_view {
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float tan_fovx, tan_fovy;
    float limx_neg, limx_pos, limy_neg, limy_pos; /* camera.h:63-66 */
    int width, height;
} glic_view;

/* ---------------------------------------------------------------------------------------
 * Workspace sizes.  Replaces required<GeometryState/ImageState/BinningState/SampleState>()
 * (rasterizer/cuda_rasterizer/rasterizer_impl.h:92-108, rasterizer_impl.cu:233-291).
 * The four buffers are opaque and may be round-tripped by the caller between forward and
 * backward (rasterizer.cpp:83-98 saves them in the autograd context).
 * ------------------------------------------------------------------------------------- */
GLIC_API size_t glic_geom_bytes(int P);
GLIC_API size_t glic_image_bytes(int width, int height);
GLIC_API size_t glic_binning_bytes(int64_t num_rendered);
GLIC_API size_t glic_sample_bytes(int64_t num_rendered, int width, int height); /* upper bound on buckets */
GLIC_API int64_t glic_max_buckets(int64_t num_rendered, int width, int height);

/* ---------------------------------------------------------------------------------------
 * Forward, stage 1: per-Gaussian preprocess + tile counting + prefix sum.
 * Replaces FORWARD::preprocess + cub::DeviceScan::InclusiveSum + the D2H read of
 * num_rendered (rasterizer_impl.cu:362-398; forward.cu:232-319).
 *   means3D[P,3] scales[P,3] rotations[P,4] opacities[P] dc[P,3] sh[P,M,3]: activated inputs
 *   (rasterize_points.h:25-50).  sh may be NULL when M == 0.
 *   radii[P] (int32) is written for every Gaussian (0 = culled).
 *   *num_rendered_host receives R (the call synchronises `stream` once, exactly where the
 *   reference does its blocking cudaMemcpy at rasterizer_impl.cu:398).
 * ------------------------------------------------------------------------------------- */
GLIC_API int glic_forward_preprocess(int P, int sh_degree, int M, const float* means3D, const float* scales,
                            float scale_modifier, const float* rotations, const float* opacities,
                            const float* dc, const float* sh, const glic_view* view, int no_color,
                            int* radii, void* geom_ws, size_t geom_bytes, void* image_ws,
                            size_t image_bytes, int64_t* num_rendered_host, void* stream);

/* Forward, stage 2: key emission, stable radix sort of (tile|depth) keys, tile ranges, bucket
 * offsets, per-tile front-to-back blend.  Replaces duplicateWithKeys, cub::DeviceRadixSort,
 * identifyTileRanges, perTileBucketCount, the second InclusiveSum + D2H read, FORWARD::render
 * and the D2D copy of the colour image (rasterizer_impl.cu:400-473; forward.cu:321-481).
 *   out_color[3,H,W], out_final_T[H,W]; no background term (forward.cu:466-467).
 *   *num_buckets_host receives B when non-NULL (costs one stream sync; pass NULL to stay
 *   asynchronous -- backward only needs the value stored inside image_ws).
 * ------------------------------------------------------------------------------------- */
GLIC_API int glic_forward_render(int P, const glic_view* view, int no_color, int64_t num_rendered,
                        void* geom_ws, void* image_ws, void* binning_ws, size_t binning_bytes,
                        void* sample_ws, size_t sample_bytes, float* out_color, float* out_final_T,
                        int64_t* num_buckets_host, void* stream);

/* ---------------------------------------------------------------------------------------
 * Forward in ONE asynchronous call (no host synchronisation inside): the same stages as
 * glic_forward_preprocess + glic_forward_render, but the binning / sample workspaces are sized by
 * the CALLER'S CAPACITY instead of the exact num_rendered, which therefore never has to travel
 * to the host mid-frame (the reference blocks twice per forward, rasterizer_impl.cu:398,442).
 *   capacity = glic_binning_capacity(binning_bytes, sample_bytes, W, H, no_color) pairs.
 *   counters_host[3] = {R, B, overflow} is written ASYNCHRONOUSLY (use pinned memory); it is valid
 *   once `stream` has been synchronised.  overflow != 0 means R exceeded the capacity: the
 *   images are invalid and the call must be repeated with workspaces for at least R pairs.
 *   For glic_backward pass num_rendered = that same capacity (the value the workspaces were
 *   carved with), not R.
 * ------------------------------------------------------------------------------------- */
GLIC_API int64_t glic_binning_capacity(size_t binning_bytes, size_t sample_bytes, int width, int height, int no_color);
GLIC_API int glic_forward(int P, int sh_degree, int M, const float* means3D, const float* scales, float scale_modifier,
                          const float* rotations, const float* opacities, const float* dc, const float* sh,
                          const glic_view* view, int no_color, int* radii, void* geom_ws, size_t geom_bytes,
                          void* image_ws, size_t image_bytes, void* binning_ws, size_t binning_bytes, void* sample_ws,
                          size_t sample_bytes, float* out_color, float* out_final_T, int64_t* counters_host,
                          void* stream);

/* Backward.  Replaces CudaRasterizer::Rasterizer::backward (rasterizer_impl.cu:476-580;
 * backward.cu:138-597) AND the ten torch::zeros of rasterize_points.cu:192-201: every output
 * element is written (exact zeros for culled Gaussians), nothing needs pre-zeroing.
 *   dL_dpix[3,H,W]; outputs shaped as RasterizeGaussiansBackwardCUDA's tensors
 *   (rasterize_points.h:52-82): dL_dmeans2D[P,3] (NDC-scaled, .z = 0), dL_dcolors[P,3],
 *   dL_dopacity[P], dL_dmeans3D[P,3], dL_dcov3D[P,6], dL_ddc[P,3], dL_dsh[P,M,3],
 *   dL_dscales[P,3], dL_drotations[P,4].  dL_dconic[P,4] (x,y,0,w) is scratch and may alias
 *   nothing else.  Gradients are w.r.t. the ACTIVATED inputs.
 * ------------------------------------------------------------------------------------- */
GLIC_API int glic_backward(int P, int sh_degree, int M, const float* means3D, const float* scales,
                  float scale_modifier, const float* rotations, const float* dc, const float* sh,
                  const glic_view* view, const int* radii, int64_t num_rendered, const void* geom_ws,
                  const void* binning_ws, const void* image_ws, const void* sample_ws,
                  const float* dL_dpix, float lambda_erank, float* dL_dmeans2D, float* dL_dconic,
                  float* dL_dopacity, float* dL_dcolors, float* dL_dmeans3D, float* dL_dcov3D,
                  float* dL_ddc, float* dL_dsh, float* dL_dscales, float* dL_drotations, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stable LSD radix sort of (uint64 key, uint32 value) pairs on key bits [0, end_bit).
 * Replaces cub::DeviceRadixSort::SortPairs<uint64,uint32>(begin_bit=0, end_bit)
 * (rasterizer_impl.cu:419-424).  Own onesweep implementation, no CUB.
 * Result lands in keys_out/vals_out; keys_in/vals_in are clobbered (used as ping-pong).
 * ------------------------------------------------------------------------------------- */
GLIC_API size_t glic_sort_temp_bytes(int64_t n);
GLIC_API int glic_sort_pairs_u64_u32(int64_t n, int end_bit, uint64_t* keys_in, uint32_t* vals_in,
                            uint64_t* keys_out, uint32_t* vals_out, void* temp, size_t temp_bytes,
                            void* stream);

/* ---------------------------------------------------------------------------------------
 * Visibility-masked Adam.  Replaces ADAM::adamUpdate (rasterizer/cuda_rasterizer/adam.cu:9-67,
 * adam.h:12-23).  In place on param / exp_avg / exp_avg_sq; element j belongs to Gaussian j/M;
 * no bias correction.  `visible` is N bytes (bool).
 * ------------------------------------------------------------------------------------- */
GLIC_API int glic_adam_update(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                     const uint8_t* visible, float lr, float b1, float b2, float eps, uint32_t N,
                     uint32_t M, void* stream);

/* ---------------------------------------------------------------------------------------
 * Packed model step (extension, no single reference symbol).  The model lives in one planar buffer
 *     rotation[4P] | xyz[3P] | log-scale[3P] | opacity logit[P] | dc[3P] | sh-rest[3MP]      (glic_packed_offsets)
 * which is also the layout of the packed gradient buffer (glic_p2p_allreduce_mean's payload).
 *  - glic_activations_forward replaces GaussianModel::getOpacity / getScaling / getRotation (gaussian.cpp:147-175:
 *    torch::sigmoid, torch::exp, normalize) with one kernel;
 *  - glic_activations_backward applies their chain rule IN PLACE to the gradients glic_backward wrote (what autograd's
 *    SigmoidBackward / ExpBackward / normalize backward do for the reference);
 *  - glic_adam_update_packed is SparseGaussianAdam::custom_step (optim_utils.h:102-137) for all six groups in ONE
 *    launch, element-wise identical to six glic_adam_update calls; lr6_host in buffer order
 *    {rotation, xyz, scaling, opacity, f_dc, f_rest}.
 * ------------------------------------------------------------------------------------- */
GLIC_API size_t glic_packed_floats(uint32_t P, uint32_t M);
GLIC_API int glic_packed_offsets(uint32_t P, uint32_t M, size_t* offsets6_host);
GLIC_API int glic_activations_forward(int P, const float* opacity_logit, const float* log_scale, const float* rot_raw,
                                      float* opacity, float* scale, float* rot, void* stream);
GLIC_API int glic_activations_backward(int P, const float* opacity, const float* scale, const float* rot_raw,
                                       float* dL_dopacity, float* dL_dscale, float* dL_drot, void* stream);
GLIC_API int glic_adam_update_packed(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                     const uint8_t* visible, const float* lr6_host, float b1, float b2, float eps,
                                     uint32_t P, uint32_t M, void* stream);

/* ---------------------------------------------------------------------------------------
 * extend() on the GPU (gaussian.cpp:499-638): which LiDAR points of the newest frame become Gaussians
 * and with which initial parameters.  One 64-bit atomicMin per in-image point on a (orderable depth | index) word per
 * pixel replaces the reference's host-side unordered_map<std::string, ...>; survivors (nearest point of their pixel,
 * depth_rsp > 0, rendered alpha < 0.99) are compacted in ascending index order and initialised.  Pinned against the CPU
 * oracle's restatement (tests/test_gpu_extend.py).  Device pointers except R_cw_host[9] (row-major), t_cw_host[3] and
 * count_host.  final_T comes from a no_color forward of the newest view.  glic_mapper_extend wires it to the model.
 * ------------------------------------------------------------------------------------- */
GLIC_API size_t glic_extend_bytes(int n, int width, int height);
GLIC_API int glic_extend(int n, const float* points, const float* colors, const float* depth_rsp, const float* R_cw_host,
                         const float* t_cw_host, float fx, float fy, float cx, float cy, int width, int height,
                         const float* final_T, float scaling_scale, void* ws, size_t ws_bytes, int* keep_idx, float* xyz,
                         float* f_dc, float* log_scale, float* rot, float* opacity_logit, int* count_host, void* stream);

/* ---------------------------------------------------------------------------------------
 * On-disk map.  Replaces GaussianModel::saveMap (gaussian.cpp:305-397, through tinyply.h:588-703):
 * binary little-endian PLY, element vertex, float properties x y z | f_dc_0..2 | f_rest_0..3M-1 | opacity | scale_0..2 |
 * rot_0..3; RAW parameters; f_rest channel-major.  The file is byte-identical to the reference's for the same tensors
 * (pinned against the reference's own tinyply in tests/test_map_io.py).  All pointers are HOST pointers; f_dc is
 * [P,1,3], f_rest [P,M,3] in the in-memory (coefficient-major) layout.  glic_ply_read* accept exactly this layout
 * (comments allowed) and fail loudly on anything else.
 * ------------------------------------------------------------------------------------- */
GLIC_API size_t glic_ply_bytes(uint32_t P, uint32_t M);
GLIC_API int glic_ply_write(const char* path, uint32_t P, uint32_t M, const float* xyz, const float* f_dc, const float* f_rest,
                            const float* opacity, const float* scale, const float* rotation);
GLIC_API int glic_ply_write_packed(const char* path, uint32_t P, uint32_t M, const float* params_host);
GLIC_API int glic_ply_read_header(const char* path, uint32_t* P, uint32_t* M, size_t* data_offset);
GLIC_API int glic_ply_read(const char* path, uint32_t P, uint32_t M, float* xyz, float* f_dc, float* f_rest, float* opacity,
                           float* scale, float* rotation);

/* ---------------------------------------------------------------------------------------
 * Fused SSIM.  Replaces fusedssimCUDA / fusedssim_backwardCUDA (fused-ssim/ssim.cu:186-365).
 * img*: [B,CH,H,W] contiguous.  Partial-derivative maps may be NULL in forward (train = false).
 * ------------------------------------------------------------------------------------- */
GLIC_API int glic_fused_ssim(int B, int CH, int H, int W, float C1, float C2, const float* img1, const float* img2,
                    float* ssim_map, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12, void* stream);
GLIC_API int glic_fused_ssim_backward(int B, int CH, int H, int W, float C1, float C2, const float* img1,
                             const float* img2, const float* dL_dmap, const float* dm_dmu1,
                             const float* dm_dsigma1_sq, const float* dm_dsigma12, float* dL_dimg1,
                             void* stream);

/* Whole photometric loss of one mapping iteration in two launches (gaussian.cpp:685-691,
 * loss_utils.h:30-33,189-193): L = (1-lambda)*mean|img-gt| + lambda*(1-mean(SSIM)), and
 * dL/dimg.  loss_out: 1 device float (accumulated; zeroed by the call).  scratch: 3*CH*H*W floats. */
GLIC_API size_t glic_loss_scratch_bytes(int CH, int H, int W);
GLIC_API int glic_l1_ssim_loss(int CH, int H, int W, float lambda_dssim, const float* img, const float* gt,
                      float* loss_out, float* dL_dimg, void* scratch, size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * simple-knn.  Replaces SimpleKNN::knn (simple-knn/simple_knn.cu:185-221): mean of the three
 * smallest squared distances to other points.  points[P,3] -> mean_dists[P].
 * ------------------------------------------------------------------------------------- */
GLIC_API size_t glic_knn_temp_bytes(int P);
GLIC_API int glic_knn_mean_dist2(int P, const float* points, float* mean_dists, void* temp, size_t temp_bytes,
                        void* stream);

/* ---------------------------------------------------------------------------------------
 * Introspection for parity tests (copies internal state to caller DEVICE buffers; any pointer
 * may be NULL).  Shapes: depth[P] xy[P,2] conic_opacity[P,4] rgb[P,3] tiles_touched[P]
 * offsets[P] (end of each Gaussian's slot range in the depth-ordered key list) clamped[P,3] (u8) |
 * point_list[R] keys_sorted[R] ((tile<<32)|depth bits, rebuilt) | ranges[T,2] bucket_offsets[T]
 * n_contrib[H*W] max_contrib[T].
 * ------------------------------------------------------------------------------------- */
GLIC_API int glic_debug_geom(int P, const void* geom_ws, float* depth, float* xy, float* conic_opacity, float* rgb,
                    uint32_t* tiles_touched, uint32_t* offsets, uint8_t* clamped, void* stream);
/* num_rendered = the value the binning workspace was carved with (exact R, or the capacity after glic_forward);
 * count = how many sorted entries to copy out (the true R). */
GLIC_API int glic_debug_binning(int P, const void* geom_ws, int64_t num_rendered, int64_t count, const void* binning_ws,
                                uint32_t* point_list, uint64_t* keys_sorted, void* stream);
GLIC_API int glic_debug_image(int width, int height, const void* image_ws, uint32_t* ranges, uint32_t* bucket_offsets,
                     uint32_t* n_contrib, uint32_t* max_contrib, int64_t* counters2_host /* {R,B}, HOST pointer */, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU exchange step (no counterpart in the reference, which is single-GPU): in-place MEAN all-reduce of the
 * packed per-Gaussian gradients + OR of the visibility bytes, done by our own kernels over NVLink peer memory
 * (CUDA IPC mappings of every rank's buffer; two-shot: rank r reduces slice r from all peers and writes it back
 * to all peers).  Buffer layout: n_floats fp32 | n_vis_bytes bytes | flag words, each part padded to 256 B
 * (glic_p2p_buffer_bytes).  Every rank must make the same sequence of calls (the barrier epoch is kept on the device,
 * so the call is CUDA-graph capturable).
 * glic_p2p_alloc is the one place this library allocates device memory: IPC needs a dedicated cudaMalloc block.
 * ------------------------------------------------------------------------------------- */
GLIC_API size_t glic_p2p_buffer_bytes(size_t n_floats, size_t n_vis_bytes);
GLIC_API int glic_p2p_slice(int rank, int world, size_t n_floats, size_t n_vis_bytes, size_t* out6_host);
GLIC_API int glic_p2p_alloc(size_t bytes, void** dev_ptr_host, unsigned char* handle64_host);
GLIC_API int glic_p2p_open(const unsigned char* handle64_host, void** peer_ptr_host);
GLIC_API int glic_p2p_close(void* peer_ptr);
GLIC_API int glic_p2p_free(void* dev_ptr);
GLIC_API int glic_p2p_allreduce_mean(int rank, int world, void* const* bufs_host, size_t n_floats, size_t n_vis_bytes,
                                     void* stream);
/* The device-side barriers wait at most GLIC_P2P_TIMEOUT_MS (environment, default 20000) for a peer; a timeout sets a
 * sticky error word in the rank's own block instead of hanging or trapping.  glic_p2p_check synchronises `stream` and
 * returns GLIC_ERR_TIMEOUT if one was recorded since the last check (and clears it). */
GLIC_API int glic_p2p_check(void* own_buf, size_t n_floats, size_t n_vis_bytes, void* stream);

/* Fused exchange + optimiser on the FULL packed gradient form: reduce-scatter of the gradients, visibility-masked Adam on
 * the local slice, all-gather of the updated PARAMETERS, in one kernel over peer memory (same NVLink bytes as the all-reduce,
 * Adam traffic and moment buffers / world).  Buffer = glic_p2p_model_bytes: gradients | visibility | flags | parameters
 * (packed layout).  Bit-identical to all-reduce + glic_adam_update_packed (tests/test_gpu_p2p_adam.py).  The mapper uses the
 * compact exchange instead (colour-gradient push + 11-float reduce, see "Native mapping host"). */
GLIC_API size_t glic_p2p_model_bytes(size_t n_floats, size_t n_vis_bytes);
GLIC_API int glic_p2p_reduce_adam(int rank, int world, void* const* bufs_host, uint32_t P, uint32_t M, float* exp_avg,
                                  float* exp_avg_sq, const float* lr6_host, float b1, float b2, float eps, void* stream);

/* ---------------------------------------------------------------------------------------
 * Evaluation metrics (evaluateVisualQuality, gaussian.cpp:721-830 with loss_utils.h:35-39,
 * 84-127): both images are clamped to [0,1]; PSNR = 10 log10(1 / mean((a-b)^2)); SSIM = mean of the 11x11 /
 * sigma 1.5 zero-padded SSIM map (the conv2d formulation, same window as the fused kernel).  out2[0] = PSNR,
 * out2[1] = SSIM (device floats).  scratch: glic_eval_scratch_bytes.  LPIPS stays TorchScript (out of scope).
 * ------------------------------------------------------------------------------------- */
GLIC_API size_t glic_eval_scratch_bytes(int CH, int H, int W);
GLIC_API int glic_eval_psnr_ssim(int CH, int H, int W, const float* img, const float* gt, float* out2, void* scratch,
                                 size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Capacity-laid-out model arena (replaces the torch::cat re-allocation of all parameters and Adam
 * state in GaussianModel::densificationPostfix, gaussian.cpp:426-497).  Same planar group order as the packed model,
 * but the group offsets come from the CAPACITY:  rotation[4*Pcap] | xyz[3*Pcap] | log-scale[3*Pcap] | opacity[Pcap] |
 * dc[3*Pcap] | sh-rest[3M*Pcap]  (glic_packed_offsets(Pcap, M)); rows [0, P) of every group are live.  Appending n rows
 * writes rows [P, P+n) in place (Adam moments of the new rows = 0, sh-rest = 0, like the zeros_like extension tensors);
 * growing copies the six live prefixes into an arena of a larger capacity.  Pcap must be a multiple of 4.
 * ------------------------------------------------------------------------------------- */
GLIC_API int glic_arena_append(float* params, float* exp_avg, float* exp_avg_sq, uint32_t P, uint32_t Pcap, uint32_t M,
                               uint32_t n_new, const float* xyz, const float* f_dc, const float* log_scale, const float* rot,
                               const float* opacity_logit, void* stream);
GLIC_API int glic_arena_regrow(const float* src, uint32_t Pcap_src, float* dst, uint32_t Pcap_dst, uint32_t M, uint32_t P,
                               void* stream);

/* ---------------------------------------------------------------------------------------
 * Native mapping host: the C++ counterpart of GaussianModel + extend() + optimize() +
 * evaluateVisualQuality() + saveMap() (gaussian.cpp:113-830) on top of the kernels above -- one object per GPU, no
 * torch, no Python.  The mapper OWNS its device memory (arena, Adam state, workspaces, keyframe images stay pinned on
 * the host like Camera::original_image_, gaussian.cpp:80): it is the host of the loop, not an operator.
 *
 *   create -> initialize(raw parameters) -> { add_keyframe -> [extend] -> optimize }* -> evaluate / save_map
 *
 * One iteration (gaussian.cpp:674-716) = for each of this rank's views: H2D of the pinned image (copy stream,
 * double-buffered), fused activations, forward (capacity-sized binning, no host sync), fused L1/D-SSIM loss, backward
 * with the activation chain rule fused in; then ONE masked Adam launch.  Gradients never exist as [P,59]: the
 * backward leaves 11 geometric floats (rotation, xyz, log-scale, opacity logit; summed over the rank's views) plus the
 * clamp-masked dL/dcolour (3 floats) of every view, and the Adam kernel rebuilds dL/d(dc, sh-rest) -- linear in
 * dL/dcolour with a per-view SH basis -- on the fly.  With world > 1 (view-sharded data parallelism) the
 * exchange step moves exactly that compact form over NVLink peer memory: colour gradients are pushed to every peer while
 * the per-Gaussian backward still runs, the 11 geometric floats are mean-reduced by the two-shot kernel; every replica
 * then applies the identical Adam step (replicas stay bit-identical).  Semantics: mean of the per-view losses over all
 * world * views_per_rank views, one optimiser step per batch.
 * All host pointers unless stated.  Status codes as everywhere; the mapper never throws.
 * ------------------------------------------------------------------------------------- */
typedef struct glic_mapper glic_mapper;

typedef struct glic_mapper_config {
    int width, height;
    float fx, fy, cx, cy;          /* pinhole intrinsics of every keyframe (mapping.h:53-118 Params) */
    int sh_degree;                 /* 0..3; M = (sh_degree+1)^2 - 1 */
    float position_lr, feature_lr, opacity_lr, scaling_lr, rotation_lr;  /* config/fastlivo.yaml:18-22; sh-rest: feature_lr/20 */
    float lambda_dssim;            /* 0.2 */
    float scaling_scale;           /* extend(): log(scaling_scale * depth / focal), gaussian.cpp:623 */
    uint32_t capacity;             /* initial arena capacity in Gaussians (doubles when an append does not fit; fixed when world > 1) */
    int max_iters;                 /* optimize(): 100 (gaussian.cpp:645) */
    uint64_t seed;                 /* view sampler (the reference seeds from std::random_device) */
    int rank, world;               /* world > 1: call glic_mapper_export / glic_mapper_connect before the first iteration */
    int views_per_rank;            /* views each rank renders per iteration (gradient accumulation); >= 1 */
} glic_mapper_config;

typedef struct glic_keyframe {     /* Camera (camera.h:38-110): pose of the camera in the world + its image */
    float R_wc[9];                 /* row-major */
    float t_wc[3];
    const float* image;            /* [3,H,W] in [0,1], host memory that stays valid (pinned => asynchronous H2D) */
} glic_keyframe;

typedef struct glic_mapper_stats {
    uint32_t num_gaussians, capacity;
    uint64_t iterations;           /* optimiser steps so far */
    uint32_t last_inserted;        /* Gaussians appended by the last extend() */
    float last_loss;               /* loss of the last finished iteration's last local view */
    double mean_visible;           /* optimize()'s return value: mean #visible Gaussians per iteration (gaussian.cpp:719) */
    uint32_t overflow_regrows;     /* times the binning capacity had to grow (the frame is redone) */
    float ms_extend, ms_optimize;  /* device time of the last extend() / optimize() */
} glic_mapper_stats;

GLIC_API int glic_mapper_create(const glic_mapper_config* cfg, glic_mapper** out);
GLIC_API int glic_mapper_destroy(glic_mapper* m);
/* raw parameters (host): xyz[P,3] f_dc[P,3] f_rest[P,M,3] (may be NULL: zeros) opacity_logit[P] log_scale[P,3] rot[P,4] */
GLIC_API int glic_mapper_initialize(glic_mapper* m, uint32_t P, const float* xyz, const float* f_dc, const float* f_rest,
                                    const float* opacity_logit, const float* log_scale, const float* rot);
GLIC_API int glic_mapper_add_keyframe(glic_mapper* m, const glic_keyframe* kf, int is_train);
/* extend() (gaussian.cpp:499-638) on the newest training keyframe: alpha-only render, z-buffer de-duplication of the n
 * LiDAR points, filter, initialisation, in-place append.  points[n,3] colors[n,3] depth_rsp[n] are host arrays. */
GLIC_API int glic_mapper_extend(glic_mapper* m, int n, const float* points, const float* colors, const float* depth_rsp);
/* optimize() (gaussian.cpp:640-719): view sampler (all keyframes when <= max_iters, else a random subset; shuffled),
 * then one iteration per sampled view batch.  views_host (optional, n_views entries) overrides the sampler. */
GLIC_API int glic_mapper_optimize(glic_mapper* m, const int* views_host, int n_views);
/* the sampler alone (tests): writes up to max_iters keyframe indices, returns their number in *count */
GLIC_API int glic_mapper_sample_views(glic_mapper* m, int* views_host, int* count);
/* PSNR / SSIM of one keyframe (train = 1: training list, 0: held-out list); synchronises */
GLIC_API int glic_mapper_evaluate(glic_mapper* m, int is_train, int index, float* psnr, float* ssim);
GLIC_API int glic_mapper_save_map(glic_mapper* m, const char* path);
GLIC_API int glic_mapper_stats_get(glic_mapper* m, glic_mapper_stats* out);
/* live rows of the arena copied to host arrays shaped like glic_mapper_initialize's (any pointer may be NULL); synchronises */
GLIC_API int glic_mapper_download(glic_mapper* m, float* xyz, float* f_dc, float* f_rest, float* opacity_logit,
                                  float* log_scale, float* rot, float* exp_avg_packed, float* exp_avg_sq_packed);
/* multi-GPU wiring: every rank exports the 64-byte IPC handle of its exchange block, the caller all-gathers them
 * (torch.distributed / MPI / a file: plumbing) and hands all `world` handles back. */
GLIC_API int glic_mapper_export(glic_mapper* m, unsigned char* handle64);
GLIC_API int glic_mapper_connect(glic_mapper* m, const unsigned char* handles /* [world][64] */);
GLIC_API int glic_mapper_synchronize(glic_mapper* m);
/* options: GLIC_MAPPER_OPT_OPTIMIZER (default 1): 0 leaves the Adam launch out of the iteration, so that a benchmark can time
 * the rasterization step (activations, forward, loss, backward, exchange) on its own.  Keyframe images may be device pointers
 * (the copy uses cudaMemcpyDefault): that is how "inputs resident in HBM" is measured. */
enum { GLIC_MAPPER_OPT_OPTIMIZER = 1, GLIC_MAPPER_OPT_BINNING_PAIRS = 2 /* (tile, Gaussian) pairs the binning workspace holds; grows by itself on overflow */ };
GLIC_API int glic_mapper_set_option(glic_mapper* m, int option, int value);
/* Camera (camera.h:38-110) in closed form, as the mapper derives it from a keyframe pose: out41 = view[16] (column-major
 * Rt) | proj[16] (column-major P*Rt) | campos[3] | tan_fovx, tan_fovy | limx_neg, limx_pos, limy_neg, limy_pos. */
GLIC_API int glic_camera_block(int width, int height, float fx, float fy, float cx, float cy, const float* R_wc /*[9] row-major*/,
                               const float* t_wc /*[3]*/, float* out41);

/* ---------------------------------------------------------------------------------------
 * Stage timing (cudaEvent pairs recorded on the launching stream around each stage; off by
 * default).  glic_profile_enable(1) starts recording and resets the accumulators;
 * glic_profile_read() synchronises the recorded events and returns, per stage, the summed
 * milliseconds and the number of recordings since the last enable.  Stage ids: GLIC_STAGE_*.
 * ------------------------------------------------------------------------------------- */
enum {
    GLIC_STAGE_PREPROCESS = 0, GLIC_STAGE_EMIT = 1, GLIC_STAGE_SORT = 2, GLIC_STAGE_RANGES = 3,
    GLIC_STAGE_RENDER_FWD = 4, GLIC_STAGE_LOSS_FWD = 5, GLIC_STAGE_LOSS_BWD = 6, GLIC_STAGE_RENDER_BWD = 7,
    GLIC_STAGE_PREPROCESS_BWD = 8, GLIC_STAGE_ADAM = 9, GLIC_STAGE_ZERO = 10, GLIC_STAGE_ALLREDUCE = 11, GLIC_STAGE_COUNT = 12
};
GLIC_API int glic_profile_enable(int on);
GLIC_API int glic_profile_read(float* ms_host /*[GLIC_STAGE_COUNT]*/, int* count_host /*[GLIC_STAGE_COUNT]*/);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
GLIC_API uint64_t glic_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GLIC_B200_H_INCLUDED */
