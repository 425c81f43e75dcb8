// voxe_query.cu -- stand-alone point queries of the voxel grid (sm_100a): VoxelGrid.forward / forward_attn of the
// reference (thre3d_atom/thre3d_reprs/voxels.py:287-345, 347-406) outside the ray-marcher.
//
// The reference answers a query of N points with two grid_sample calls (densities, after the pre-activation has been
// applied to the WHOLE grid; features) and a concatenation: [N, F + 1] = (interpolated features, post(interpolated
// pre(density * scale))), zeros padding outside the grid, no inside mask (the mask belongs to the renderer,
// renderers.py:81-86).  Here one thread answers one 16-byte channel group of one point from the packed volume the
// render kernels read: 8 vector loads, 8 FMAs per channel, a coalesced row write.  The backward scatters the upstream
// rows into the packed gradient volume with vector reductions, exactly like the ray-marcher's backward.
//
// Unlike a render sample, a query point may lie anywhere.  The packed volume only carries a one-voxel apron, so a point
// whose footprint leaves the aproned extent on some axis (floor(u) outside [0, N]) is answered without touching
// memory: every corner it could address is padding, the interpolated values are zero and the density is post(0).
#include "voxe_device.cuh"
#include "voxe_launch.h"

namespace voxe {
namespace {

struct QueryPoint {
  Corners c;
  bool covered;  // the footprint addresses the aproned volume (NaN coordinates count as covered and propagate)
};

__device__ __forceinline__ bool axis_covered(float pc, float ua, float ub, int N) {
  const float fl = floorf(fmaf(pc, ua, ub));
  return !(fl < 0.f || fl > (float)N);
}

__device__ __forceinline__ void locate(const KParams& p, const float* __restrict__ points, long long pt, QueryPoint& q) {
  const float px = __ldg(points + 3 * pt), py = __ldg(points + 3 * pt + 1), pz = __ldg(points + 3 * pt + 2);
  q.covered = axis_covered(px, p.ua[0], p.ub[0], p.X) & axis_covered(py, p.ua[1], p.ub[1], p.Y) & axis_covered(pz, p.ua[2], p.ub[2], p.Z);
  make_corners(p, px, py, pz, q.c);  // indices are clamped into the volume whatever the point
}

// Interpolated pre-activated density from the 8 corner vectors that hold the density channel, and one bit per corner
// whose pre-activation derivative is negative (abs only).
__device__ __forceinline__ float density_raw(const KParams& p, const Corners& c, const float4 (&v)[8], int dco, unsigned& signs) {
  float sig = 0.f;
  signs = 0u;
  if (p.preact == kPreAbs) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float dv = f4_get(v[q], dco);
      if (dv * p.dscale < 0.f) signs |= (1u << q);
      sig = fmaf(c.w[q], fabsf(dv), sig);
    }
    return sig * fabsf(p.dscale);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) sig = fmaf(c.w[q], f4_get(v[q], dco), sig);
  return sig * p.dscale;
}

// out [N, F + 1]; thread = (point, 16-byte channel group j of cv)
__global__ void __launch_bounds__(256) query_points_kernel(const KParams p, const float* __restrict__ points, float* __restrict__ out,
                                                           long long n, int cv, int F) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * cv) return;
  const long long pt = i / cv;
  const int j = (int)(i - pt * cv);
  QueryPoint q;
  locate(p, points, pt, q);
  float4 v[8];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q.covered) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(p.grid + (size_t)q.c.idx[k] * cv + j);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float w = q.c.w[k];
      acc.x = fmaf(w, v[k].x, acc.x);
      acc.y = fmaf(w, v[k].y, acc.y);
      acc.z = fmaf(w, v[k].z, acc.z);
      acc.w = fmaf(w, v[k].w, acc.w);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float* row = out + pt * (F + 1);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = 4 * j + k;
    if (ch < F) {
      row[ch] = f4_get(acc, k);
    } else if (ch == F) {
      unsigned signs;
      float dpost;
      row[ch] = post_act(p.postact, density_raw(p, q.c, v, k, signs), dpost);
    }
  }
}

// g_out [N, F + 1] -> packed gradient volume (+=)
__global__ void __launch_bounds__(256) query_points_bwd_kernel(const KParams p, const float* __restrict__ points,
                                                               const float* __restrict__ g_out, long long n, int cv, int F) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * cv) return;
  const long long pt = i / cv;
  const int j = (int)(i - pt * cv);
  QueryPoint q;
  locate(p, points, pt, q);
  if (!q.covered) return;
  const float* row = g_out + pt * (F + 1);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  int dco = -1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = 4 * j + k;
    if (ch < F) f4_set(g, k, __ldg(row + ch));
    else if (ch == F) dco = k;
  }
  unsigned signs = 0u;
  float gden = 0.f;  // dL / d(interpolated pre-activated density) x d(pre-activated density)/d(stored density), sign aside
  if (dco >= 0) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(p.grid + (size_t)q.c.idx[k] * cv + j);
    float dpost;
    post_act(p.postact, density_raw(p, q.c, v, dco, signs), dpost);
    gden = __ldg(row + F) * dpost * (p.preact == kPreAbs ? fabsf(p.dscale) : p.dscale);
  }
  if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f && gden == 0.f) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float w = q.c.w[k];
    float4 t = make_float4(w * g.x, w * g.y, w * g.z, w * g.w);
    if (dco >= 0) f4_set(t, dco, ((signs >> k) & 1u) ? -w * gden : w * gden);
    red_add_v4(p.grad + (size_t)q.c.idx[k] * cv + j, t.x, t.y, t.z, t.w);
  }
}

}  // namespace

cudaError_t launch_query_points(const KParams& p, const float* points, float* out, const float* g_out, long long n, int cv,
                                int n_features, cudaStream_t stream) {
  const int threads = 256;
  const long long blocks = (n * cv + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  if (g_out == nullptr)
    query_points_kernel<<<(unsigned)blocks, threads, 0, stream>>>(p, points, out, n, cv, n_features);
  else
    query_points_bwd_kernel<<<(unsigned)blocks, threads, 0, stream>>>(p, points, g_out, n, cv, n_features);
  return cudaGetLastError();
}

}  // namespace voxe
