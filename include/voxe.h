/*
 * voxe.h -- C ABI of libvoxe_sm100a.so: the B200-native replacement for Vox-E's differentiable
 *           SH-voxel-grid ray-marcher (the hot path behind `render_sh_voxel_grid`).
 *
 * Reference interface each entry point replaces (paths relative to the Vox-E tree):
 *
 *   voxe_render_fwd   the three stages bound by  thre3d_atom/thre3d_reprs/renderers.py:50-105  and run by
 *                     thre3d_atom/rendering/volumetric/render_interface.py:140-171 :
 *                       sampler      thre3d_atom/rendering/volumetric/sample.py:15-68, 71-184, 187-202
 *                       processor    thre3d_atom/rendering/volumetric/process.py:20-96  (VoxelGrid.forward,
 *                                    thre3d_atom/thre3d_reprs/voxels.py:287-342; SH ladder
 *                                    thre3d_atom/rendering/volumetric/utils/spherical_harmonics.py:64-132)
 *                       accumulator  thre3d_atom/rendering/volumetric/accumulate.py:31-113
 *   voxe_render_bwd   the autograd backward of the above, triggered by `total_loss.backward()` at
 *                     thre3d_atom/modules/trainers.py:350 and thre3d_atom/modules/sds_trainer.py:332
 *   voxe_pack_grid /  the per-call full-grid passes of VoxelGrid.forward (voxels.py:303-305: `_densities *
 *   voxe_unpack_grad  expected_density_scale`, pre-activation) and their backward: the kernels work on ONE packed
 *                     channel-last volume [X,Y,Z,C] (C = features + density, padded to a multiple of 4) so a
 *                     trilinear corner is one or more 16-byte vectors and a gradient scatter is a vector RED.
 *   attn mode         VOXE_FLAG_ATTN selects the twin  renderers.py:108-163 / accumulate.py:115-198  (one colour
 *                     channel, background term forced to zero).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data owned by the caller (torch allocations in the Python
 *     binding); the library allocates nothing persistent and keeps no global state except a thread-local error
 *     string, the optional tuning override and the run-time handle of libnccl (voxe_allreduce_grads only);
 *   - kernels are enqueued asynchronously on `stream`; the calls never synchronise;
 *   - return value 0 = success, otherwise a VOXE_ERR_* code (or a cudaError_t offset by VOXE_ERR_CUDA_BASE);
 *     `voxe_last_error()` describes the most recent failure on the calling thread;
 *   - there is no CPU fallback: the calls fail when no sm_100-class device/kernel image is available.
 */
#ifndef VOXE_H_
#define VOXE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOXE_ABI_VERSION 14

#if defined(__GNUC__)
#define VOXE_API __attribute__((visibility("default")))
#else
#define VOXE_API
#endif

/* opaque CUDA stream handle (binary compatible with cudaStream_t / CUstream) */
typedef struct CUstream_st* voxe_stream_t;

/* density pre-activation (applied to `densities * density_scale` at the voxels, voxels.py:303-305) */
enum { VOXE_PREACT_IDENTITY = 0, VOXE_PREACT_ABS = 1 };
/* density post-activation (applied to the interpolated density, voxels.py:320) */
enum { VOXE_POSTACT_IDENTITY = 0, VOXE_POSTACT_RELU = 1, VOXE_POSTACT_SOFTPLUS = 2 };

/* render flags == the boolean fields of SHVoxGridRenderConfig (renderers.py:29-47) */
enum {
  VOXE_FLAG_PERTURB = 1,           /* perturb_sampled_points: stratified jitter, needs `jitter` [R,S]      */
  VOXE_FLAG_AABB_SAMPLING = 2,     /* optimized_sampling: per-ray slab test, misses fall back to near/far  */
  VOXE_FLAG_DISPARITY_SAMPLING = 4,/* linear_disparity_sampling (ignored with AABB sampling, as upstream)  */
  VOXE_FLAG_WHITE_BKGD = 8,        /* white_bkgd                                                           */
  VOXE_FLAG_RENDER_DIFFUSE = 16,   /* render_diffuse: only the degree-0 SH coefficient                     */
  VOXE_FLAG_ATTN = 32              /* attention-grid twin: n_colour == 1, no background term               */
};

enum {
  VOXE_OK = 0,
  VOXE_ERR_INVALID_ARGUMENT = 1,
  VOXE_ERR_UNSUPPORTED = 2,       /* combination outside the fused set (e.g. SH degree > 3, S > 4096)       */
  VOXE_ERR_NO_DEVICE = 3,
  VOXE_ERR_CUDA_BASE = 1000,      /* + cudaError_t */
  VOXE_ERR_NCCL_BASE = 2000       /* + ncclResult_t (voxe_allreduce_grads and the voxe_nccl_* helpers) */
};

/* Geometry + activations of a voxel grid (VoxelGrid, voxels.py:46-130). */
typedef struct VoxeGridDesc {
  int32_t dims[3];        /* X, Y, Z (width_x, depth_y, height_z); z is the fastest-varying axis            */
  int32_t n_features;     /* F = n_colour * (sh_degree+1)^2                                                 */
  int32_t channels;       /* packed channels per voxel: voxe_packed_channels(F) = roundup4(F + 1)           */
  float aabb_lo[3];       /* fp32(grid_location - dims*voxel_size/2)   (voxels.py:198-223)                  */
  float aabb_hi[3];
  float norm_scale[3];    /* fp32(2)/(fp32(hi)-fp32(lo))               (imaging_utils.py:57-63)             */
  float norm_bias[3];     /* fp32(-1) - fp32(lo)*norm_scale                                                 */
  float density_scale;    /* expected_density_scale                                                         */
  int32_t preact;         /* VOXE_PREACT_*                                                                  */
  int32_t postact;        /* VOXE_POSTACT_*                                                                 */
} VoxeGridDesc;

/* One render call (SHVoxGridRenderConfig, renderers.py:29-47). */
typedef struct VoxeRenderDesc {
  int32_t num_samples;    /* S >= 2                                                                         */
  float near;             /* camera_bounds.near                                                             */
  float far;              /* camera_bounds.far                                                              */
  int32_t flags;          /* VOXE_FLAG_*                                                                    */
  int32_t sh_degree;      /* 0..3                                                                           */
  int32_t n_colour;       /* 3 (RGB) or 1 (attn)                                                            */
  float noise_std;        /* stochastic_density_noise_std; != 0 needs `noise` [R,S] (standard normal)       */
  uint64_t rng_seed;      /* VOXE_FLAG_PERTURB with jitter == NULL: the U[0,1) draws of sample.py:63 are generated */
  uint64_t rng_offset;    /* inside the kernels: a counter-based PCG hash of (rng_seed, rng_offset, ray, sample)     */
  const int64_t* rng_seed_dev;   /* NULL, or DEVICE pointers to the seed and the offset of a generator whose state is  */
  const int64_t* rng_offset_dev; /* registered with a CUDA graph (torch's PhiloxCudaState under capture): the kernels   */
  uint64_t rng_offset_intragraph;/* then use (*rng_seed_dev, *rng_offset_dev + rng_offset_intragraph), re-read on replay */
  uint64_t* stats;        /* NULL, or four DEVICE counters the backward adds to (measurement runs): [0] += in-grid    */
                          /* samples it processed, [1] += samples whose 8-corner scatter it issued (a sample whose    */
                          /* gradient is exactly zero -- e.g. ReLU at a non-positive density -- scatters nothing),    */
                          /* [2] += scattering samples whose cell no lower lane of the same warp instruction hits,    */
                          /* [3] += corner REDs whose voxel no lower lane hits (what a perfect intra-warp merge keeps) */
} VoxeRenderDesc;

VOXE_API int voxe_abi_version(void);
VOXE_API const char* voxe_last_error(void);

/* roundup4(n_features + 1): channel count of the packed volume. */
VOXE_API int voxe_packed_channels(int n_features);

/* Number of floats of the packed volume of `grid`: channels * 8 * ceil((X+2)/2) * ceil((Y+2)/2) * ceil((Z+2)/2).  The
 * volume is the grid plus a one-voxel apron of zeros (grid_sample's zeros padding, voxels.py:306-333, without range
 * checks in the kernels), stored as 2x2x2 bricks of voxels (one brick of SH-0 voxels = one 128-byte line) -- an internal
 * layout: only voxe_pack_grid / voxe_unpack_grad convert to and from the reference's [X,Y,Z,.] tensors. */
VOXE_API int64_t voxe_packed_floats(const VoxeGridDesc* grid);

/* packed <- bricked concat(features[X,Y,Z,F], densities[X,Y,Z,1], zero padding).  Values are copied verbatim;
 * density scale and pre-activation are applied inside the render kernels at gather time. */
VOXE_API int voxe_pack_grid(const VoxeGridDesc* grid, const float* densities, const float* features, float* packed,
                   voxe_stream_t stream);

/* Split a packed gradient volume into d_densities[X,Y,Z,1] and d_features[X,Y,Z,F].
 * accumulate != 0: add into the outputs; == 0: overwrite.  Either output may be NULL (skipped). */
VOXE_API int voxe_unpack_grad(const VoxeGridDesc* grid, const float* packed_grad, float* d_densities, float* d_features,
                     int accumulate, voxe_stream_t stream);

/* Number of floats of the `saved` workspace that links a forward call to its backward call: one 16-byte vector per
 * sample slot (tone-mapped colour(s) + interpolated raw density, so the backward does not re-gather the 8 corner
 * vectors of every sample) plus (n_colour + 3) floats per ray and depth segment (the transmittance at the segment
 * start and the segment's local sums of w*colour, w*z, w).  About 4.4 floats per sample, against the ~35 floats per
 * sample autograd keeps for the reference.  The buffer must be 16-byte aligned.  Depends on the current
 * voxe_set_tuning() state: do not retune between the two calls. */
VOXE_API int64_t voxe_saved_floats(const VoxeRenderDesc* render, int64_t num_rays);

/* Forward render of R rays.
 *   packed   voxe_packed_floats() floats from voxe_pack_grid
 *   rays_o/d [R,3]       origins / (un-normalised) directions
 *   jitter   [R,S] or NULL   the U[0,1) draws of sample.py:63; NULL with VOXE_FLAG_PERTURB: drawn in-kernel from
 *                            (render->rng_seed, render->rng_offset) -- pass the same pair to the backward call
 *   noise    [R,S] or NULL   the N(0,1) draws of accumulate.py:59-62 (required when noise_std != 0)
 *   colour   [R,n_colour], depth [R], acc [R], disparity [R]   outputs (disparity may be NULL)
 *   saved    voxe_saved_floats() floats, or NULL when no backward will follow (inference) */
VOXE_API int voxe_render_fwd(const VoxeGridDesc* grid, const VoxeRenderDesc* render, const float* packed,
                    const float* rays_o, const float* rays_d, const float* jitter, const float* noise,
                    float* colour, float* depth, float* acc, float* disparity, float* saved, int64_t num_rays,
                    voxe_stream_t stream);

/* Backward: per sample, reloads the 16-byte vector the forward saved (no re-gather of the 8 corners, no O(R*S)
 * activations beyond that vector), recomputes the compositing weights and scatter-ADDS dL/d(packed) into `packed_grad`
 * (same layout and size as `packed`; caller-zeroed; accumulate-into, so several ray batches or both renders of a training
 * step can share one buffer).  `saved` is the workspace the matching forward call filled (same grid, rays, jitter,
 * noise, same voxe_set_tuning state).  g_depth / g_acc / g_disp may be NULL (treated as zero).  The disparity gradient
 * is applied only on rays whose disparity is finite (depth/acc > 1e-10).
 *   touched / touch_tag   NULL, or voxe_touched_bytes() bytes (one per 2x2x2 brick of the packed volume): every sample
 *                         that scatters stores (uint8_t)touch_tag at the brick of its first corner, so that
 *                         voxe_consume_grad can visit only what this call wrote.  Flags are never cleared by the
 *                         library: use a new tag (1..255) for every backward/consume pair and memset the flags to 0
 *                         before a tag value is used again. */
VOXE_API int voxe_render_bwd(const VoxeGridDesc* grid, const VoxeRenderDesc* render, const float* packed,
                    const float* rays_o, const float* rays_d, const float* jitter, const float* noise,
                    const float* saved, const float* g_colour, const float* g_depth, const float* g_acc,
                    const float* g_disp, float* packed_grad, uint8_t* touched, int32_t touch_tag, int64_t num_rays,
                    voxe_stream_t stream);

/* Bytes of a `touched` flag array for `grid`: one per 2x2x2 brick of the packed volume. */
VOXE_API int64_t voxe_touched_bytes(const VoxeGridDesc* grid);

/* Pinhole camera of voxe_render_camera (CameraIntrinsics + CameraPose, utils/imaging_utils.py:17-30). */
typedef struct VoxeCameraDesc {
  int32_t height, width;
  float focal;
  float rotation[9];      /* camera-to-world rotation, row-major                                            */
  float translation[3];   /* camera position                                                                */
} VoxeCameraDesc;

/* Whole-camera inference render (no backward follows): pixels [first_pixel, first_pixel + num_pixels) in flat
 * row-major order (index = y * W + x, as flatten_rays(cast_rays(...)) orders them).  Replaces, for
 * VolumetricModel.render (modules/volumetric_model.py:135-193), cast_rays (rendering/volumetric/utils/misc.py:12-50),
 * the 32768-ray chunk loop and the per-chunk render: rays are generated inside the kernel, one thread per pixel walks
 * its ray front to back and stops once the transmittance drops below `min_transmittance` (early termination: changes a
 * pixel by less than that value; pass 0 to evaluate every sample).  Outputs as voxe_render_fwd; stratified jitter, when
 * VOXE_FLAG_PERTURB is set, comes from (render->rng_seed, render->rng_offset); density noise is not supported here. */
VOXE_API int voxe_render_camera(const VoxeGridDesc* grid, const VoxeRenderDesc* render, const VoxeCameraDesc* camera,
                       const float* packed, int64_t first_pixel, int64_t num_pixels, float* colour, float* depth,
                       float* acc, float* disparity, float min_transmittance, voxe_stream_t stream);

/* The same inference kernel on caller-supplied rays [R,3] (forward only, one thread per ray, early termination): for
 * callers that already hold ray tensors, and for AABB-bound sampling (VOXE_FLAG_AABB_SAMPLING), where the first / last
 * sample sits exactly on a grid face and the picture depends on the last bit of the ray direction -- the rays must then
 * be the very tensors cast_rays produced.  `jitter` [R,S] or NULL as in voxe_render_fwd. */
VOXE_API int voxe_render_infer(const VoxeGridDesc* grid, const VoxeRenderDesc* render, const float* packed,
                      const float* rays_o, const float* rays_d, const float* jitter, float* colour, float* depth,
                      float* acc, float* disparity, int64_t num_rays, float min_transmittance, voxe_stream_t stream);

/* The jitter the kernels generate for (render->rng_seed, render->rng_offset): out[R, S], for tests and replays. */
VOXE_API int voxe_jitter_fill(const VoxeRenderDesc* render, float* out, int64_t num_rays, voxe_stream_t stream);

/* Sparse hand-over of a packed gradient volume: every non-zero 16-byte vector of `packed_grad` is ADDED into
 * d_densities[X,Y,Z,1] / d_features[X,Y,Z,F] (either may be NULL) and then cleared, so that `packed_grad` is all-zero
 * again when the call returns.  A ray batch touches a few percent of the voxels; this pass reads the volume once and
 * writes only what the batch touched, replacing zero-fill + voxe_unpack_grad + a dense `.grad +=` per backward call
 * (the autograd accumulation of trainers.py:350 / sds_trainer.py:332).  Requires d_* to hold valid values (zeros for
 * a fresh gradient).
 * With `touched` (the flag array the backward call(s) since the last consume wrote with `touch_tag`) only bricks that can
 * hold gradients are read: ~1 byte per brick instead of the whole volume (0.5 MB against 68 MB at 160^3). */
VOXE_API int voxe_consume_grad(const VoxeGridDesc* grid, float* packed_grad, float* d_densities, float* d_features,
                      const uint8_t* touched, int32_t touch_tag, voxe_stream_t stream);

/* One Adam step on the two grids, fused with everything else the optimiser step does to them
 * (replaces `optimizer.zero_grad(); ...; optimizer.step()` of thre3d_atom/modules/trainers.py:247-255,348-351 and
 * thre3d_atom/modules/sds_trainer.py:198-203,332-334 for torch.optim.Adam(betas, eps), no weight decay / amsgrad):
 *   gradient = packed_grad (what voxe_render_bwd accumulated; may be NULL) + dense_d_* (reference-layout gradients of
 *   torch-side losses such as the TV / density-correlation terms of sds_trainer.py:290-326; may be NULL);
 *   `packed` must mirror the parameters (voxe_pack_grid) and is updated in place, the new values are also written to
 *   densities / features (reference layout), the moments packed_m / packed_v live in the packed layout
 *   (voxe_packed_floats() floats each; convert with voxe_pack_grid / voxe_unpack_grad for checkpoints), and
 *   `packed_grad` is zeroed -- one streaming pass, ten 16-byte-coalesced streams.
 * `step` is the 1-based step count AFTER this update (bias corrections 1 - beta^step).  `densities` or `features` may be
 * NULL: that tensor is frozen (requires_grad == False upstream) -- its values and moments are left untouched and whatever
 * the backward scattered for it is discarded. */
typedef struct VoxeAdamDesc {
  double lr;     /* doubles, like the Python scalars torch.optim.Adam derives 1 - beta and lr / (1 - beta1^step) from */
  double beta1;
  double beta2;
  double eps;
  int32_t step;
} VoxeAdamDesc;

VOXE_API int voxe_adam_step(const VoxeGridDesc* grid, const VoxeAdamDesc* adam, float* densities, float* features,
                            float* packed, float* packed_grad, const float* dense_d_densities,
                            const float* dense_d_features, float* packed_m, float* packed_v, voxe_stream_t stream);

/* Rescale a channel-last grid [X,Y,Z,C] to [X2,Y2,Z2,C] covering the same world extent -- the
 * `interpolate(mode="trilinear", align_corners=False, size=output_size)` of scale_voxel_grid_with_required_output_size
 * (thre3d_atom/thre3d_reprs/voxels.py:409-447, called between the stages of progressive training, trainers.py:155 and :481)
 * without the concatenate / permute / slice copies around it: call it once per tensor (features, densities, attn). */
VOXE_API int voxe_resample_grid(const float* grid_in, const int32_t in_dims[3], int32_t channels, float* grid_out,
                                const int32_t out_dims[3], voxe_stream_t stream);

/* Stand-alone point queries: VoxelGrid.forward / forward_attn (thre3d_atom/thre3d_reprs/voxels.py:287-345, 347-406) outside
 * the ray-marcher.  `points` [N,3] world coordinates, anywhere (zeros padding outside the grid, NO inside mask -- that one
 * belongs to the renderer); `out` [N, n_features + 1] = (interpolated features, post(interpolated pre(density * scale))),
 * read from the packed volume the render kernels use (pack the attention grid as `features` for forward_attn).
 * voxe_query_points_bwd adds the gradient of sum(out * g_out) to the packed gradient volume (convert with
 * voxe_unpack_grad); points receive no gradient (the reference's callers never differentiate them). */
VOXE_API int voxe_query_points(const VoxeGridDesc* grid, const float* packed, const float* points, float* out, int64_t num_points,
                               voxe_stream_t stream);
VOXE_API int voxe_query_points_bwd(const VoxeGridDesc* grid, const float* packed, const float* points, const float* g_out,
                                   float* packed_grad, int64_t num_points, voxe_stream_t stream);

/* ---- per-step full-grid regularisers of the edit loop --------------------------------------------
 * `workspace` is VOXE_REG_WORKSPACE_DOUBLES doubles of device memory owned by the caller, ZEROED ONCE when it is allocated
 * (per-CTA partial sums of the reductions -- no floating-point atomics, so a loss is bitwise reproducible --, a ticket by
 * which the last CTA of a launch folds them, left at zero again by every call, and, for the pair loss, the statistics its
 * gradient call reads back); calls that share a workspace must be ordered on one stream; `loss` is ONE device float.  `upstream` (one device float,
 * dL/dloss as autograd hands it over; NULL = 1) and the host scalar `scale` (a loss weight) multiply the gradient, which
 * is ADDED into `grad` when accumulate != 0 and overwrites it otherwise. */
#define VOXE_REG_WORKSPACE_DOUBLES 8192

/* Total-variation loss of a channel-last grid [X,Y,Z,C] and/or its gradient, one streaming pass:
 *   loss = (mean|diff_x h| + mean|diff_y h| + mean|diff_z h|) / 3,  h = relu ? max(grid, 0) : grid
 * -- `_tv_loss_on_grid` (thre3d_atom/modules/sds_trainer.py:563-567, attn_grid_trainer.py:659-663, grid_refine.py:709-713)
 * as applied to ReLU(_densities) and _features at sds_trainer.py:318-326 and to the attention grids at
 * attn_grid_trainer.py:361-366.  `loss` NULL: gradient only; `grad` NULL: loss only; both: one pass for both.  An axis of
 * extent 1 makes the loss NaN (torch's mean over an empty tensor) and contributes no gradient. */
VOXE_API int voxe_tv_regularizer(const float* grid, const int32_t dims[3], int32_t channels, int32_t relu, double* workspace,
                                 float* loss, const float* upstream, float scale, float* grad, int32_t accumulate,
                                 voxe_stream_t stream);

/* Loss between the edited density grid `a` and the frozen pretrained one `b` (n floats each):
 *   VOXE_PAIR_CORRELATION  1 - mean((a-mean a)(b-mean b)) / (sqrt(var a * var b) + 1e-7)   `_density_correlation_loss`,
 *                          sds_trainer.py:507-524; `correlation_grid` (n floats or NULL) receives its second return value
 *   VOXE_PAIR_L2 / _L1     mse_loss / l1_loss(a, b)                                        sds_trainer.py:498-503
 * voxe_pair_loss_grad writes dloss/da; in correlation mode it reads the statistics voxe_pair_loss left in `workspace`
 * (same a, b), so call it after voxe_pair_loss on the same stream. */
enum { VOXE_PAIR_CORRELATION = 0, VOXE_PAIR_L2 = 1, VOXE_PAIR_L1 = 2 };
VOXE_API int voxe_pair_loss(const float* a, const float* b, int64_t n, int32_t mode, double* workspace, float* loss,
                            float* correlation_grid, voxe_stream_t stream);
VOXE_API int voxe_pair_loss_grad(const float* a, const float* b, int64_t n, int32_t mode, const double* workspace,
                                 const float* upstream, float scale, float* grad, int32_t accumulate, voxe_stream_t stream);

/* ---- training-side ray-batch sampling ---------------------------------------------------------------
 * Replaces `sample_random_rays_and_pixels_synchronously` (thre3d_atom/rendering/volumetric/utils/misc.py:126-138, called
 * every iteration at thre3d_atom/modules/trainers.py:311) and the per-view `cast_rays` + `collate_rays` that feed it
 * (trainers.py:290-301; misc.py:12-50): `sample_size` threads each take the i-th element of a keyed pseudo-random
 * permutation of [0, num_pixels) -- distinct indices, like `torch.randperm(N)[:sample_size]`, without the O(N) shuffle --
 * and produce that pixel's ray and colour.
 *   camera mode (poses != NULL)   poses [B,3,4] = [R | t] per image (the dataset's pose matrices, trainers.py:293-295);
 *                                 index = (b*H + row)*W + col; the ray is generated as cast_rays does; no ray tensors exist
 *   gather mode (poses == NULL)   rays are gathered from src_rays_o / src_rays_d [num_pixels,3] (the reference signature)
 *   pixels [num_pixels, C] -> pixels_out [sample_size, C] (both may be NULL)
 *   indices_in [sample_size] or NULL: use these indices instead of drawing (replays, parity tests);
 *   indices_out [sample_size] or NULL: the indices used;  rays_o / rays_d [sample_size,3] (may be NULL: indices only). */
typedef struct VoxeSamplerDesc {
  int64_t num_pixels;
  int32_t height, width;   /* camera mode only */
  float focal;
  int32_t pixel_channels;
  uint64_t rng_seed;       /* the permutation is a function of (rng_seed, rng_offset, num_pixels) only */
  uint64_t rng_offset;
} VoxeSamplerDesc;

VOXE_API int voxe_sample_rays(const VoxeSamplerDesc* sampler, const float* poses, const float* src_rays_o,
                              const float* src_rays_d, const float* pixels, const int64_t* indices_in, int64_t sample_size,
                              int64_t* indices_out, float* rays_o, float* rays_d, float* pixels_out, voxe_stream_t stream);

/* ---- the step's one collective (SURVEY.md row e): SUM of the packed gradient volume over the data-parallel ranks ----------
 * The reference is single-GPU; with ray batches / views sharded over ranks and the grid replicated, the dense gradients
 * autograd would have produced for the whole batch (trainers.py:350, sds_trainer.py:332) are the sum of the ranks' volumes.
 *
 * voxe_allreduce_grads_peer: own two-shot kernel over peer-mapped memory (NVLink / NVSwitch), in place, one launch per
 * rank on `stream`; all ranks must call it with the same n_floats (a multiple of 4).  Every rank's volume and signal pad
 * must be mapped into the calling process (CUDA IPC, VMM export, or torch's symmetric memory in the Python binding):
 *   buffers[k]  rank k's gradient volume (buffers[rank] is the local one), 16-byte aligned, same size everywhere
 *   signals[k]  rank k's signal pad: VOXE_SIGNAL_WORDS uint32, zeroed once at allocation, used only by this call
 *   multicast   NVLS multicast mapping of the volumes (multimem.ld_reduce / multimem.st: the switch adds and
 *               replicates), or NULL: plain peer loads and stores; multicast_share splits a launch between the two
 *   fail_flag   NULL, or a device uint32 that is OR-ed with 1 when a peer did not arrive within ~2 s (the kernel then
 *               gives up instead of hanging the GPU; the volume contents are undefined). */
#define VOXE_MAX_PEERS 16
#define VOXE_SIGNAL_WORDS 4096
typedef struct VoxePeerDesc {
  int32_t world_size, rank;
  float* buffers[VOXE_MAX_PEERS];
  uint32_t* signals[VOXE_MAX_PEERS];
  float* multicast;
  int32_t multicast_share;   /* 1..8: of every 8 CTAs, how many take the multicast path while the others use plain peer  */
                             /* loads / stores (the two are bound by different resources); 0 or 8: all (with `multicast`) */
} VoxePeerDesc;
VOXE_API int voxe_allreduce_grads_peer(const VoxePeerDesc* peers, int64_t n_floats, uint32_t* fail_flag, voxe_stream_t stream);

/* voxe_allreduce_grads_peer_sparse: the same exchange restricted to the 2x2x2 bricks that SOME rank's backward touched in
 * this step -- for volumes of which a step writes a small part (a 65 536-ray batch through a 512^3 SH-2 grid touches a few
 * percent of 15 GB).  `peers->buffers` are the ranks' PACKED gradient volumes of `grid` (voxe_packed_floats floats each);
 * `touched_peers[k]` is rank k's flag array as written by voxe_render_bwd(..., touched, touch_tag, ...), peer-mapped like the
 * volumes, 16-byte aligned and voxe_peer_touched_bytes(grid) long (voxe_touched_bytes rounded up to whole 16-byte vectors,
 * the padding zeroed once).  Two launches on `stream`: the flag arrays are all-reduced with "some rank carries the tag"
 * as the operator (afterwards every rank's array holds the union, so voxe_consume_grad(..., touched, touch_tag) visits
 * exactly the bricks that hold sums), then the tagged bricks are summed over the ranks in place.  Same tag on every rank. */
VOXE_API int64_t voxe_peer_touched_bytes(const VoxeGridDesc* grid);
VOXE_API int voxe_allreduce_grads_peer_sparse(const VoxePeerDesc* peers, const VoxeGridDesc* grid, uint8_t* const* touched_peers,
                                              int32_t touch_tag, uint32_t* fail_flag, voxe_stream_t stream);

/* voxe_allreduce_grads: ncclAllReduce(SUM, fp32, in place) of `buf` on `nccl_comm` (an ncclComm_t) -- for hosts whose
 * volumes are not peer-mapped.  libnccl.so.2 is opened at run time (no link-time dependency); the helpers below build a
 * communicator without any other framework: rank 0 calls voxe_nccl_unique_id and ships the VOXE_NCCL_UNIQUE_ID_BYTES bytes
 * to the other ranks by its own means, then every rank calls voxe_nccl_comm_create (collective). */
#define VOXE_NCCL_UNIQUE_ID_BYTES 128
VOXE_API int voxe_nccl_unique_id(void* id_out);
VOXE_API int voxe_nccl_comm_create(void** nccl_comm_out, int32_t world_size, int32_t rank, const void* id);
VOXE_API int voxe_nccl_comm_destroy(void* nccl_comm);
VOXE_API int voxe_allreduce_grads(void* nccl_comm, float* buf, int64_t n_floats, voxe_stream_t stream);

/* Launch-shape override for tuning runs: samples per thread (1..64; the number of depth segments per ray is
 * ceil(S / samples_per_thread)), rays per CTA (1..32) and the register budget of the kernel variant
 * (64, 80, 96 or 128); 0 restores the built-in choice of that knob.  Does not change results beyond fp32 summation
 * order.  The 80- and 96-register variants are limited to 128 threads per CTA. */
VOXE_API int voxe_set_tuning(int samples_per_thread, int rays_per_cta, int register_cap);

/* Number of kernels this library has launched on the calling process since load (for bench accounting). */
VOXE_API int64_t voxe_launch_count(void);

/* Render launches that took a flag-specialised kernel variant (same arithmetic as the generic kernels with the per-call
 * switches resolved at compile time; the default wherever a variant exists, environment VOXE_SPECIALISED_KERNELS=0, read
 * once at the first render call, forces the generic kernels). */
VOXE_API int64_t voxe_specialised_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VOXE_H_ */
