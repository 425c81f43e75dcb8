// voxe_grid_ops.cu -- full-grid streaming passes around the ray-marcher (sm_100a).
//
// The reference keeps `_densities [X,Y,Z,1]` and `_features [X,Y,Z,F]` as two parameters (voxels.py:98-114) and
// VoxelGrid.forward touches the whole density grid on every call (voxels.py:303-305).  The render kernels instead
// read ONE packed channel-last volume so that a trilinear corner is a 16-byte vector; these two kernels convert
// between the layouts.  They are pure HBM streams: 2*(F+1)*4 bytes per voxel each.
#include <cmath>

#include "voxe_launch.h"

namespace voxe {
namespace {

// Packed volume layout: the grid plus a one-voxel apron of zeros on every face ((X+2) x (Y+2) x (Z+2) voxels, index
// shifted by +1), cut into 2x2x2 bricks, bricks in (x, y, z) row-major order, 8 voxel slots per brick ordered
// (x&1, y&1, z&1), CV float4 per slot.  At SH-0 (CV = 1) a brick is exactly one 128-byte line, so the 8 corners of a
// trilinear cell touch 1..8 lines (3.4 on average) instead of 4..8, and neighbouring rays share lines in x and y as
// well as in z.  The apron implements grid_sample's zeros padding without range checks in the render kernels: apron
// slots (and the padding slots of partial bricks) hold zeros in the volume, and whatever the backward scatters into
// them in the gradient volume is dropped here.
struct BrickDims {
  int X, Y, Z, BY, BZ;
};

__device__ __forceinline__ bool slot_to_voxel(const BrickDims& d, int64_t slot, int64_t& v) {
  const int64_t brick = slot >> 3;
  const int within = (int)(slot & 7);
  const int bz = (int)(brick % d.BZ);
  const int64_t t = brick / d.BZ;
  const int by = (int)(t % d.BY);
  const int bx = (int)(t / d.BY);
  const int x = 2 * bx + (within >> 2) - 1, y = 2 * by + ((within >> 1) & 1) - 1, z = 2 * bz + (within & 1) - 1;
  v = ((int64_t)x * d.Y + y) * d.Z + z;
  return x >= 0 && y >= 0 && z >= 0 && x < d.X && y < d.Y && z < d.Z;  // false: apron / brick padding slot
}

// one thread per packed float4: slot = t / CV, channels 4j..4j+3
__global__ void __launch_bounds__(256) pack_grid_kernel(const float* __restrict__ dens, const float* __restrict__ feat,
                                                        float4* __restrict__ packed, int64_t n_vec, int F, int CV,
                                                        BrickDims d) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_vec) return;
  const int64_t slot = t / CV;
  const int c0 = (int)(t - slot * CV) * 4;
  int64_t v;
  float out[4] = {0.f, 0.f, 0.f, 0.f};
  if (slot_to_voxel(d, slot, v)) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + k;
      out[k] = (c < F) ? __ldg(feat + v * F + c) : (c == F ? __ldg(dens + v) : 0.f);
    }
  }
  packed[t] = make_float4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(256) unpack_grad_kernel(const float4* __restrict__ pg, float* __restrict__ d_dens,
                                                          float* __restrict__ d_feat, int64_t n_vec, int F, int CV,
                                                          int accumulate, BrickDims d) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_vec) return;
  const int64_t slot = t / CV;
  const int c0 = (int)(t - slot * CV) * 4;
  int64_t v;
  if (!slot_to_voxel(d, slot, v)) return;
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

// Sparse hand-over: add the non-zero vectors of the packed gradient volume into the reference-layout gradients and
// clear them.  One thread per packed float4; a warp reads 512 contiguous bytes and writes only where a ray batch
// scattered something (a few percent of the volume), so the pass costs one read of the volume.
__global__ void __launch_bounds__(256) consume_grad_kernel(float4* __restrict__ pg, float* __restrict__ d_dens,
                                                           float* __restrict__ d_feat, int64_t n_vec, int F, int CV,
                                                           BrickDims d) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_vec) return;
  const float4 g = pg[t];
  if (g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f) return;
  pg[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t slot = t / CV;
  const int c0 = (int)(t - slot * CV) * 4;
  int64_t v;
  if (!slot_to_voxel(d, slot, v)) return;  // apron / padding slot: whatever was scattered there is dropped
  const float in[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + k;
    if (c < F) {
      if (d_feat) d_feat[v * F + c] += in[k];
    } else if (c == F) {
      if (d_dens) d_dens[v] += in[k];
    }
  }
}

// The same hand-over when the backward call left a trail: `touched` holds one byte per 2x2x2 brick, and a sample that
// scattered stored `tag` at the brick of its corner 0 -- its 8 corners lie in that brick and the +1 neighbours in x, y, z.
// A brick can therefore hold gradients only if one of the 8 flags at (bx-dx, by-dy, bz-dz) carries the tag.  One thread
// per brick: 8 byte loads (neighbouring threads read neighbouring bytes), and only tagged bricks touch their 8 * CV
// gradient vectors.  A 4096-ray batch tags a few percent of the bricks, so the pass reads ~1 byte per brick (0.5 MB at
// 160^3) instead of the whole 68 MB volume.  Flags are not cleared: the next call uses another tag, and a stale tag only
// costs a look at vectors that are zero (see voxe.h).
__global__ void __launch_bounds__(256) consume_touched_kernel(float4* __restrict__ pg, float* __restrict__ d_dens,
                                                              float* __restrict__ d_feat,
                                                              const unsigned char* __restrict__ touched, int tag,
                                                              unsigned n_bricks, int F, int CV, BrickDims d) {
  asm volatile("griddepcontrol.wait;" ::: "memory");  // chained behind the backward kernel (launch_chained): its scatter is complete from here
  const unsigned brick = blockIdx.x * blockDim.x + threadIdx.x;
  if (brick >= n_bricks) return;
  const unsigned bz = brick % (unsigned)d.BZ, t = brick / (unsigned)d.BZ;
  const unsigned by = t % (unsigned)d.BY, bx = t / (unsigned)d.BY;
  const unsigned char want = (unsigned char)tag;
  bool any = false;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const unsigned dx = k >> 2, dy = (k >> 1) & 1, dz = k & 1;
    if (bx >= dx && by >= dy && bz >= dz) any |= __ldg(touched + ((bx - dx) * (unsigned)d.BY + (by - dy)) * (unsigned)d.BZ + (bz - dz)) == want;
  }
  if (!any) return;
  float4* base = pg + (int64_t)brick * 8 * CV;
  for (int j = 0; j < CV; ++j) {
    float4 g[8];  // the j-th vector of the brick's 8 voxel slots: eight independent loads in flight, then the adds
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = base[k * CV + j];
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // voxel slot (x&1, y&1, z&1) of this brick
      if (g[k].x == 0.f && g[k].y == 0.f && g[k].z == 0.f && g[k].w == 0.f) continue;
      base[k * CV + j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int x = 2 * (int)bx + (k >> 2) - 1, y = 2 * (int)by + ((k >> 1) & 1) - 1, z = 2 * (int)bz + (k & 1) - 1;
      if (!(x >= 0 && y >= 0 && z >= 0 && x < d.X && y < d.Y && z < d.Z)) continue;  // apron / padding slot: dropped
      const int64_t v = ((int64_t)x * d.Y + y) * d.Z + z;
      const float in[4] = {g[k].x, g[k].y, g[k].z, g[k].w};
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int c = 4 * j + c4;
        if (c < F) {
          if (d_feat) atomicAdd(d_feat + v * F + c, in[c4]);  // fire-and-forget RED: no load -> add -> store round trip
        } else if (c == F) {
          if (d_dens) atomicAdd(d_dens + v, in[c4]);
        }
      }
    }
  }
}

// Multi-vector voxels (SH degree >= 1: a brick is 8 * CV vectors, 896 bytes at SH-2): one thread per brick would walk its
// brick in 112-byte strides while its neighbours do the same 896 bytes further on -- every load instruction touches 32
// lines.  Here a warp takes 32 consecutive bricks, one lane per brick for the flag test, ballots, and then all lanes
// together stream the vectors of the flagged bricks: consecutive lanes read consecutive 16-byte vectors, and the adds of a
// voxel's channels land in one contiguous row of the dense gradient.  The adds stay fire-and-forget REDs although every
// dense element belongs to exactly one thread of the launch: a plain load-add-store was measured slower (cfg 5: 1.39 ms
// against 1.20 ms; the round trip of the load costs more than the L2 atomic unit's 140 M sectors,
// profiles/r2_cfg5_handover_kernel_metrics.txt).
__global__ void __launch_bounds__(256) consume_touched_tiles_kernel(float4* __restrict__ pg, float* __restrict__ d_dens,
                                                                    float* __restrict__ d_feat,
                                                                    const unsigned char* __restrict__ touched, int tag,
                                                                    unsigned n_bricks, int F, int CV, BrickDims d) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const unsigned tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned brick = tile * 32u + lane;
  const unsigned char want = (unsigned char)tag;
  bool any = false;
  if (brick < n_bricks) {
    const unsigned bz = brick % (unsigned)d.BZ, t = brick / (unsigned)d.BZ;
    const unsigned by = t % (unsigned)d.BY, bx = t / (unsigned)d.BY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned dx = k >> 2, dy = (k >> 1) & 1, dz = k & 1;
      if (bx >= dx && by >= dy && bz >= dz) any |= __ldg(touched + ((bx - dx) * (unsigned)d.BY + (by - dy)) * (unsigned)d.BZ + (bz - dz)) == want;
    }
  }
  const unsigned mask = __ballot_sync(0xffffffffu, any);
  if (mask == 0u) return;
  const int vpb = 8 * CV;
  const int n = __popc(mask) * vpb;
  float4* tile_base = pg + (int64_t)tile * 32 * vpb;
  constexpr int U = 4;
  for (int i0 = lane; i0 < n; i0 += 32 * U) {
    float4 g[U];
    int off[U], within[U];
    unsigned which[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int id = i0 + 32 * u;
      off[u] = -1;
      if (id < n) {
        const int j = id / vpb;
        within[u] = id - j * vpb;
        which[u] = __fns(mask, 0u, j + 1);
        off[u] = (int)which[u] * vpb + within[u];
        g[u] = tile_base[off[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (off[u] < 0 || (g[u].x == 0.f && g[u].y == 0.f && g[u].z == 0.f && g[u].w == 0.f)) continue;
      tile_base[off[u]] = make_float4(0.f, 0.f, 0.f, 0.f);
      const unsigned b = tile * 32u + which[u];
      const unsigned bz = b % (unsigned)d.BZ, t = b / (unsigned)d.BZ;
      const unsigned by = t % (unsigned)d.BY, bx = t / (unsigned)d.BY;
      const int k = within[u] / CV, j = within[u] - k * CV;  // voxel slot (x&1, y&1, z&1) of the brick, vector of the voxel
      const int x = 2 * (int)bx + (k >> 2) - 1, y = 2 * (int)by + ((k >> 1) & 1) - 1, z = 2 * (int)bz + (k & 1) - 1;
      if (!(x >= 0 && y >= 0 && z >= 0 && x < d.X && y < d.Y && z < d.Z)) continue;  // apron / padding slot: dropped
      const int64_t v = ((int64_t)x * d.Y + y) * d.Z + z;
      const float in[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int c = 4 * j + c4;
        if (c < F) {
          if (d_feat) atomicAdd(d_feat + v * F + c, in[c4]);
        } else if (c == F) {
          if (d_dens) atomicAdd(d_dens + v, in[c4]);
        }
      }
    }
  }
}

// Fused per-step grid pass: consume the packed gradient volume (plus optional dense gradients from torch-side losses),
// apply one Adam step to the reference-layout parameters and their moments, refresh the packed volume and zero the
// packed gradients -- one launch instead of zero-fill + unpack + AccumulateGrad + ~10 optimiser kernels + repack.
// Arithmetic follows torch.optim.Adam (single-tensor path, amsgrad=False, weight_decay=0, maximize=False):
//   m <- m + (g - m) * (1 - beta1);  v <- v * beta2 + g * g * (1 - beta2)
//   p <- p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// One thread per packed float4 (4 channels of one voxel).  Pure HBM stream: 36 bytes per channel-slot.
struct AdamScalars {
  float one_minus_beta1, beta2, one_minus_beta2, step_size, sqrt_bc2, eps;
};

// One thread per packed float4 (4 channels of one voxel slot).  Gradients, parameters (read from the packed volume,
// which mirrors them), and both moments live in the bricked layout, so eight of the streams are perfectly coalesced
// 16-byte accesses; only the write-back of the new values into the reference-layout parameter tensors is scattered
// (a 12-byte feature row + a 4-byte density per voxel at SH-0, z-neighbours adjacent).
__global__ void __launch_bounds__(256) adam_step_kernel(float4* __restrict__ packed, const float4* __restrict__ packed_grad,
                                                        float4* __restrict__ packed_m, float4* __restrict__ packed_v,
                                                        float* __restrict__ dens, float* __restrict__ feat,
                                                        const float* __restrict__ dense_gd, const float* __restrict__ dense_gf,
                                                        int64_t n_vec, int F, int CV, AdamScalars a, BrickDims d) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_vec) return;
  const int64_t slot = t / CV;
  const int c0 = (int)(t - slot * CV) * 4;
  int64_t v;
  if (!slot_to_voxel(d, slot, v)) return;  // padding slots of partial bricks stay zero
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (packed_grad) {
    const float4 pg = packed_grad[t];
    g[0] = pg.x; g[1] = pg.y; g[2] = pg.z; g[3] = pg.w;  // zeroed after the kernel by a memset node, see launch_adam_step
  }
  const float4 p4 = packed[t], m4 = packed_m[t], v4 = packed_v[t];
  float p[4] = {p4.x, p4.y, p4.z, p4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + k;
    if (c > F) continue;  // channel padding
    if ((c < F) ? (feat == nullptr) : (dens == nullptr)) continue;  // frozen tensor: value and moments stay as they are
    float* dst = (c < F) ? feat + v * F + c : dens + v;
    float grad = g[k];
    if (c < F) {
      if (dense_gf) grad += __ldg(dense_gf + v * F + c);
    } else if (dense_gd) {
      grad += __ldg(dense_gd + v);
    }
    m[k] = m[k] + (grad - m[k]) * a.one_minus_beta1;
    vv[k] = vv[k] * a.beta2 + grad * grad * a.one_minus_beta2;
    const float denom = sqrtf(vv[k]) / a.sqrt_bc2 + a.eps;
    p[k] = p[k] - a.step_size * (m[k] / denom);
    *dst = p[k];
  }
  packed[t] = make_float4(p[0], p[1], p[2], p[3]);
  packed_m[t] = make_float4(m[0], m[1], m[2], m[3]);
  packed_v[t] = make_float4(vv[0], vv[1], vv[2], vv[3]);
}

// Progressive-training rescale of a channel-last grid [X,Y,Z,C] -> [X2,Y2,Z2,C]: what
// torch.nn.functional.interpolate(mode="trilinear", align_corners=False) with an explicit output size computes
// (scale_voxel_grid_with_required_output_size, voxels.py:409-447 upstream), without the cat / permute / slice copies around
// it.  Per axis: source coordinate s = (o + 0.5) * (in / out) - 0.5 clamped at 0, i0 = (int)s, i1 = i0 + (i0 < in-1),
// lambda1 = s - i0 -- ATen's area_pixel_compute_source_index, evaluated in fp32 like ATen's CUDA kernel, and blended in
// its order (z innermost).  One thread per output element; the channels of a voxel are neighbouring threads.
struct ResampleAxis {
  int n_in, n_out;
  float scale;
};

__device__ __forceinline__ void resample_axis(const ResampleAxis& a, int o, int& i0, int& i1, float& l0, float& l1) {
  float s = a.scale * ((float)o + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i0 = i0 < a.n_in - 1 ? i0 : a.n_in - 1;
  i1 = i0 + (i0 < a.n_in - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.0f - l1;
}

__global__ void __launch_bounds__(256) resample_grid_kernel(const float* __restrict__ in, float* __restrict__ out, int C, ResampleAxis ax,
                                                            ResampleAxis ay, ResampleAxis az, int64_t n_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_out) return;
  const int c = (int)(t % C);
  int64_t v = t / C;
  const int z = (int)(v % az.n_out);
  v /= az.n_out;
  const int y = (int)(v % ay.n_out), x = (int)(v / ay.n_out);
  int x0, x1, y0, y1, z0, z1;
  float lx0, lx1, ly0, ly1, lz0, lz1;
  resample_axis(ax, x, x0, x1, lx0, lx1);
  resample_axis(ay, y, y0, y1, ly0, ly1);
  resample_axis(az, z, z0, z1, lz0, lz1);
  auto at = [&](int xi, int yi, int zi) { return __ldg(in + (((int64_t)xi * ay.n_in + yi) * az.n_in + zi) * C + c); };
  out[t] = lx0 * (ly0 * (lz0 * at(x0, y0, z0) + lz1 * at(x0, y0, z1)) + ly1 * (lz0 * at(x0, y1, z0) + lz1 * at(x0, y1, z1))) +
           lx1 * (ly0 * (lz0 * at(x1, y0, z0) + lz1 * at(x1, y0, z1)) + ly1 * (lz0 * at(x1, y1, z0) + lz1 * at(x1, y1, z1)));
}

BrickDims brick_dims(const int dims[3]) { return BrickDims{dims[0], dims[1], dims[2], (dims[1] + 3) / 2, (dims[2] + 3) / 2}; }

}  // namespace

int64_t packed_voxel_slots(const int dims[3]) {
  return (int64_t)((dims[0] + 3) / 2) * ((dims[1] + 3) / 2) * ((dims[2] + 3) / 2) * 8;  // N + 2 voxels per axis (apron)
}

cudaError_t launch_pack_grid(const float* densities, const float* features, float* packed, const int dims[3],
                             int n_features, int channels, cudaStream_t stream) {
  const int CV = channels / 4;
  const int64_t n_vec = packed_voxel_slots(dims) * CV;
  const int threads = 256;
  const int64_t blocks = (n_vec + threads - 1) / threads;
  pack_grid_kernel<<<(unsigned)blocks, threads, 0, stream>>>(densities, features, reinterpret_cast<float4*>(packed),
                                                             n_vec, n_features, CV, brick_dims(dims));
  return cudaGetLastError();
}

cudaError_t launch_unpack_grad(const float* packed_grad, float* d_densities, float* d_features, const int dims[3],
                               int n_features, int channels, bool accumulate, cudaStream_t stream) {
  const int CV = channels / 4;
  const int64_t n_vec = packed_voxel_slots(dims) * CV;
  const int threads = 256;
  const int64_t blocks = (n_vec + threads - 1) / threads;
  unpack_grad_kernel<<<(unsigned)blocks, threads, 0, stream>>>(reinterpret_cast<const float4*>(packed_grad),
                                                               d_densities, d_features, n_vec, n_features, CV,
                                                               accumulate ? 1 : 0, brick_dims(dims));
  return cudaGetLastError();
}

int64_t packed_bricks(const int dims[3]) { return packed_voxel_slots(dims) / 8; }

cudaError_t launch_consume_grad(float* packed_grad, float* d_densities, float* d_features, const int dims[3],
                                int n_features, int channels, const unsigned char* touched, int tag, cudaStream_t stream) {
  const int CV = channels / 4;
  const int64_t n_vec = packed_voxel_slots(dims) * CV;
  const int threads = 256;
  if (touched != nullptr) {
    const int64_t n_bricks = packed_voxel_slots(dims) / 8;  // < 2^28 (check_grid bounds the vector count by 2^31)
    if (CV > 1)  // one lane per brick for the flags, whole warps for the flagged bricks' vectors
      return launch_chained(consume_touched_tiles_kernel, dim3((unsigned)((n_bricks + threads - 1) / threads)), dim3(threads), 0, stream,
                            reinterpret_cast<float4*>(packed_grad), d_densities, d_features, touched, tag, (unsigned)n_bricks, n_features,
                            CV, brick_dims(dims));
    return launch_chained(consume_touched_kernel, dim3((unsigned)((n_bricks + threads - 1) / threads)), dim3(threads), 0, stream,
                          reinterpret_cast<float4*>(packed_grad), d_densities, d_features, touched, tag, (unsigned)n_bricks, n_features,
                          CV, brick_dims(dims));
  }
  const int64_t blocks = (n_vec + threads - 1) / threads;
  consume_grad_kernel<<<(unsigned)blocks, threads, 0, stream>>>(reinterpret_cast<float4*>(packed_grad), d_densities,
                                                                d_features, n_vec, n_features, CV, brick_dims(dims));
  return cudaGetLastError();
}

cudaError_t launch_resample_grid(const float* in, const int in_dims[3], int channels, float* out, const int out_dims[3],
                                 cudaStream_t stream) {
  ResampleAxis a[3];
  for (int k = 0; k < 3; ++k) a[k] = ResampleAxis{in_dims[k], out_dims[k], (float)in_dims[k] / (float)out_dims[k]};
  const int64_t n_out = (int64_t)out_dims[0] * out_dims[1] * out_dims[2] * channels;
  const int threads = 256;
  resample_grid_kernel<<<(unsigned)((n_out + threads - 1) / threads), threads, 0, stream>>>(in, out, channels, a[0], a[1], a[2], n_out);
  return cudaGetLastError();
}

cudaError_t launch_adam_step(float* packed, float* packed_grad, float* packed_m, float* packed_v, float* densities,
                             float* features, const float* dense_gd, const float* dense_gf, const int dims[3], int n_features,
                             int channels, double lr, double beta1, double beta2, double eps, int step, cudaStream_t stream) {
  const int CV = channels / 4;
  const int64_t n_vec = packed_voxel_slots(dims) * CV;
  // bias corrections in double, like the Python scalars of torch.optim.Adam
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  AdamScalars a;
  a.one_minus_beta1 = (float)(1.0 - beta1);
  a.beta2 = (float)beta2;
  a.one_minus_beta2 = (float)(1.0 - beta2);
  a.step_size = (float)(lr / bc1);
  a.sqrt_bc2 = (float)sqrt(bc2);
  a.eps = (float)eps;
  const int threads = 256;
  const int64_t blocks = (n_vec + threads - 1) / threads;
  adam_step_kernel<<<(unsigned)blocks, threads, 0, stream>>>(
      reinterpret_cast<float4*>(packed), reinterpret_cast<const float4*>(packed_grad), reinterpret_cast<float4*>(packed_m),
      reinterpret_cast<float4*>(packed_v), densities, features, dense_gd, dense_gf, n_vec, n_features, CV, a, brick_dims(dims));
  cudaError_t e = cudaGetLastError();
  // The gradient volume is cleared with a memset, NOT from inside the kernel: on B200 a kernel that stores a uniform
  // value over a whole buffer runs ~4x slower than the same kernel storing varying data (tools/adam_micro.cu: 380 us vs
  // 98 us for this pass at 160^3), whereas cudaMemsetAsync clears 65.5 MB in 12 us.
  if (e == cudaSuccess && packed_grad) e = cudaMemsetAsync(packed_grad, 0, (size_t)n_vec * sizeof(float4), stream);
  return e;
}

}  // namespace voxe
