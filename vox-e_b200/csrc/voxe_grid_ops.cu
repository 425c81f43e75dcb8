// voxe_grid_ops.cu -- full-grid streaming passes around the ray-marcher (sm_100a).
//
// The reference keeps `_densities [X,Y,Z,1]` and `_features [X,Y,Z,F]` as two parameters (voxels.py:98-114) and
// VoxelGrid.forward touches the whole density grid on every call (voxels.py:303-305).  The render kernels instead
// read ONE packed channel-last volume so that a trilinear corner is a 16-byte vector; these two kernels convert
// between the layouts.  They are pure HBM streams: 2*(F+1)*4 bytes per voxel each.
#include "voxe_launch.h"

namespace voxe {
namespace {

// one thread per packed float4: packed[v][4j..4j+3]
__global__ void __launch_bounds__(256) pack_grid_kernel(const float* __restrict__ dens, const float* __restrict__ feat,
                                                        float4* __restrict__ packed, int64_t n_vec, int F, int CV) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_vec) return;
  const int64_t v = t / CV;
  const int c0 = (int)(t - v * CV) * 4;
  float out[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + k;
    out[k] = (c < F) ? __ldg(feat + v * F + c) : (c == F ? __ldg(dens + v) : 0.f);
  }
  packed[t] = make_float4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(256) unpack_grad_kernel(const float4* __restrict__ pg, float* __restrict__ d_dens,
                                                          float* __restrict__ d_feat, int64_t n_vec, int F, int CV,
                                                          int accumulate) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_vec) return;
  const int64_t v = t / CV;
  const int c0 = (int)(t - v * CV) * 4;
  const float4 g = __ldg(pg + t);
  const float in[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + k;
    if (c < F) {
      if (d_feat) {
        float* dst = d_feat + v * F + c;
        *dst = accumulate ? (*dst + in[k]) : in[k];
      }
    } else if (c == F) {
      if (d_dens) {
        float* dst = d_dens + v;
        *dst = accumulate ? (*dst + in[k]) : in[k];
      }
    }
  }
}

}  // namespace

cudaError_t launch_pack_grid(const float* densities, const float* features, float* packed, int64_t n_voxels,
                             int n_features, int channels, cudaStream_t stream) {
  const int CV = channels / 4;
  const int64_t n_vec = n_voxels * CV;
  const int threads = 256;
  const int64_t blocks = (n_vec + threads - 1) / threads;
  pack_grid_kernel<<<(unsigned)blocks, threads, 0, stream>>>(densities, features, reinterpret_cast<float4*>(packed),
                                                             n_vec, n_features, CV);
  return cudaGetLastError();
}

cudaError_t launch_unpack_grad(const float* packed_grad, float* d_densities, float* d_features, int64_t n_voxels,
                               int n_features, int channels, bool accumulate, cudaStream_t stream) {
  const int CV = channels / 4;
  const int64_t n_vec = n_voxels * CV;
  const int threads = 256;
  const int64_t blocks = (n_vec + threads - 1) / threads;
  unpack_grad_kernel<<<(unsigned)blocks, threads, 0, stream>>>(reinterpret_cast<const float4*>(packed_grad),
                                                               d_densities, d_features, n_vec, n_features, CV,
                                                               accumulate ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace voxe
