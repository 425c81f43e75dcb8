// voxe_launch.h -- host-side declarations shared by the translation units of libvoxe_sm100a.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "voxe.h"

namespace voxe {

// Launch a kernel of a chain of short dependent launches (fwd -> bwd -> hand-over -> fwd ...), optionally with the
// programmatic-serialization attribute, so that its CTAs may become resident while the previous kernel of the stream
// drains (the kernel itself blocks in griddepcontrol.wait until that kernel has completed; see pdl_wait in
// voxe_device.cuh).  Opt-in (VOXE_PDL=1): measured on the benchmark frame it moves nothing beyond run-to-run noise --
// pipelined frame 0.905 -> 0.893 ms, captured API frame 2.38 -> 2.32 ms, strictly serialised frame 1.54 -> 1.59 ms
// (profiles/r2_pdl_ab.txt) -- the ~2 us launch ramp it hides is not where a 15-25 us launch loses its time.  Works under
// stream capture (the edge becomes a programmatic dependency of the graph).
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("VOXE_PDL");
    return v != nullptr && v[0] == '1';
  }();
  return on;
}

template <class K, class... Args>
cudaError_t launch_chained(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const Args&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

struct KParams;
struct CameraParams;

// fused ray-marcher (voxe_render.cu)
cudaError_t launch_render(const KParams& p, int sh_degree, int n_colour, int regcap, bool backward, bool specialise,
                          cudaStream_t stream, bool* took_specialised);
cudaError_t launch_camera(const KParams& p, const CameraParams& cam, int sh_degree, int n_colour, cudaStream_t stream);
int max_threads_per_cta(int regcap);
cudaError_t launch_jitter_fill(int R, int S, unsigned long long seed, unsigned long long offset, const long long* seed_dev,
                               const long long* offset_dev, unsigned long long intragraph, float* out, cudaStream_t stream);
int saved_floats_per_segment(int n_colour, int samples_per_segment);  // segment summary + one float4 per sample

// full-grid passes (voxe_grid_ops.cu)
int64_t packed_voxel_slots(const int dims[3]);  // voxel slots of the 2x2x2-bricked volume (>= X*Y*Z)
cudaError_t launch_pack_grid(const float* densities, const float* features, float* packed, const int dims[3],
                             int n_features, int channels, cudaStream_t stream);
cudaError_t launch_unpack_grad(const float* packed_grad, float* d_densities, float* d_features, const int dims[3],
                               int n_features, int channels, bool accumulate, cudaStream_t stream);

// touched != nullptr: visit only the bricks whose (dilated) flag equals tag (written by the backward kernel)
cudaError_t launch_consume_grad(float* packed_grad, float* d_densities, float* d_features, const int dims[3],
                                int n_features, int channels, const unsigned char* touched, int tag, cudaStream_t stream);
int64_t packed_bricks(const int dims[3]);  // 2x2x2 bricks of the packed volume = bytes of a `touched` flag array

cudaError_t launch_resample_grid(const float* in, const int in_dims[3], int channels, float* out, const int out_dims[3],
                                 cudaStream_t stream);

// stand-alone point queries (voxe_query.cu); g_out == nullptr: forward into `out`, else backward into p.grad
cudaError_t launch_query_points(const KParams& p, const float* points, float* out, const float* g_out, long long n, int cv,
                                int n_features, cudaStream_t stream);

cudaError_t launch_adam_step(float* packed, float* packed_grad, float* packed_m, float* packed_v, float* densities,
                             float* features, const float* dense_gd, const float* dense_gf, const int dims[3], int n_features,
                             int channels, double lr, double beta1, double beta2, double eps, int step, cudaStream_t stream);

// per-step regularisers (voxe_regularizers.cu); *launches = kernels enqueued
cudaError_t launch_tv(const float* grid, const int dims[3], int channels, bool relu, double* workspace, float* loss,
                      const float* upstream, float scale, float* grad, bool accumulate, cudaStream_t stream, int* launches);
cudaError_t launch_pair_loss(const float* a, const float* b, int64_t n, int mode, float eps, double* workspace, float* loss,
                             float* corr_grid, cudaStream_t stream, int* launches);
cudaError_t launch_pair_grad(const float* a, const float* b, int64_t n, int mode, float eps, const double* workspace,
                             const float* upstream, float scale, float* grad, bool accumulate, cudaStream_t stream);

// training-side ray-batch sampling (voxe_sampler.cu)
cudaError_t launch_sample_rays(long long n, long long k, int H, int W, int C, float focal, unsigned long long seed,
                               unsigned long long offset, const float* poses, const float* src_o, const float* src_d,
                               const float* pixels, const long long* idx_in, long long* idx_out, float* rays_o, float* rays_d,
                               float* pixels_out, cudaStream_t stream);

// gradient all-reduce (voxe_collective.cu)
cudaError_t launch_allreduce_peer(const VoxePeerDesc& peers, int64_t n_floats, unsigned* fail_flag, cudaStream_t stream);
cudaError_t launch_allreduce_peer_sparse(const VoxePeerDesc& peers, unsigned char* const* touched_peers, int tag, const int dims[3],
                                         int channels, unsigned* fail_flag, cudaStream_t stream);
const char* nccl_unavailable();  // nullptr when libnccl could be opened
const char* nccl_error_string(int rc);
int nccl_unique_id(void* out128);
int nccl_comm_create(void** comm, int world, int rank, const void* id128);
int nccl_comm_destroy(void* comm);
int nccl_allreduce_sum_f32(void* comm, float* buf, size_t n, cudaStream_t stream);

}  // namespace voxe
