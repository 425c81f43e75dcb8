// voxe_sampler.cu -- training-side ray-batch sampling, sm_100a.
//
// Every iteration of the reference's reconstruction loop draws a batch with
//   permutation = torch.randperm(B*H*W); sampled = permutation[:sample_size]            (misc.py:126-138, trainers.py:311)
// over rays it has cast for ALL pixels of the loaded views (cast_rays per pose + collate: 24 bytes per pixel, 122 MB for
// 8 views of 800x800), i.e. a 5 M-element shuffle to pick 4096 rays.  Here a batch is `sample_size` threads:
//   index   i-th element of a keyed pseudo-random PERMUTATION of [0, N) evaluated on demand (cycle-walking Feistel network
//           over the next power of four >= N, PCG hash as round function): distinct indices, O(sample_size) work, no O(N) pass;
//   ray     generated from (pose of image b, intrinsics, pixel) exactly as cast_rays does (misc.py:12-50), or gathered from
//           a caller's ray tensors when those already exist (the reference signature);
//   pixel   gathered from the [N, C] pixel tensor.
// It is another realisation of "a uniformly random sample without replacement", not torch.randperm's numbers: parity of
// rays / pixels is pinned with injected indices, and the drawn indices against a CPU restatement of the permutation.
#include <cstdint>

#include "voxe_launch.h"

namespace voxe {
namespace {

__device__ __forceinline__ unsigned pcg(unsigned v) {
  const unsigned state = v * 747796405u + 2891336453u;
  const unsigned word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}

// Bijection of [0, 4^half_bits): balanced Feistel network, 6 rounds.
__device__ __forceinline__ unsigned long long feistel(unsigned long long v, int half_bits, const unsigned (&key)[6]) {
  const unsigned mask = (half_bits >= 32) ? 0xffffffffu : ((1u << half_bits) - 1u);
  unsigned l = (unsigned)(v >> half_bits) & mask, r = (unsigned)v & mask;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const unsigned t = l ^ (pcg(r ^ key[k]) & mask);
    l = r;
    r = t;
  }
  return ((unsigned long long)l << half_bits) | r;
}

struct Permutation {
  unsigned key[6];
  int half_bits;
  unsigned long long n;

  __device__ __forceinline__ void init(unsigned long long n_, unsigned long long seed, unsigned long long offset) {
    n = n_;
    half_bits = 1;
    while (half_bits < 32 && (1ull << (2 * half_bits)) < n) ++half_bits;
    unsigned k = pcg((unsigned)seed ^ pcg((unsigned)(seed >> 32)));
    k = pcg(k ^ (unsigned)offset);
    k = pcg(k ^ (unsigned)(offset >> 32));
#pragma unroll
    for (int j = 0; j < 6; ++j) key[j] = k = pcg(k + 0x9E3779B9u * (unsigned)(j + 1));
  }
  // cycle walking: the network permutes [0, 4^half_bits) with 4^half_bits < 4 n, so re-applying it until the value drops
  // below n (on average < 4 applications) restricts it to a permutation of [0, n)
  __device__ __forceinline__ unsigned long long at(unsigned long long i) const {
    unsigned long long v = feistel(i, half_bits, key);
    while (v >= n) v = feistel(v, half_bits, key);
    return v;
  }
};

struct SamplerParams {
  long long n;            // pixels to choose from (B * H * W, or rows of the ray tensors)
  long long k;            // sample size
  int H, W, C;
  float focal;
  unsigned long long seed, offset;
  const float* poses;     // [B, 3, 4] = [R | t] per image, or null: gather from rays_o / rays_d
  const float* src_o;     // [n, 3]
  const float* src_d;
  const float* pixels;    // [n, C] or null
  const long long* idx_in;  // injected indices [k] or null
  long long* idx_out;       // [k] or null
  float *rays_o, *rays_d, *pixels_out;
};

__global__ void __launch_bounds__(256) sample_rays_kernel(const SamplerParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.k) return;
  long long idx;
  if (p.idx_in != nullptr) {
    idx = p.idx_in[i];
  } else {
    Permutation perm;
    perm.init((unsigned long long)p.n, p.seed, p.offset);
    idx = (long long)perm.at((unsigned long long)i);
  }
  if (p.idx_out != nullptr) p.idx_out[i] = idx;
  if (p.rays_o != nullptr) {
    float o[3], d[3];
    if (p.poses != nullptr) {  // cast_rays (misc.py:12-50): pixel centres, dir = R ((x+.5-W/2)/f, -(y+.5-H/2)/f, -1), not normalised
      const long long per = (long long)p.H * p.W;
      const long long b = idx / per, pix = idx - b * per;
      const int row = (int)(pix / p.W), col = (int)(pix - (long long)row * p.W);
      const float* pose = p.poses + 12 * b;
      const float dx = __fdiv_rn(__fsub_rn((float)col + 0.5f, (float)p.W * 0.5f), p.focal);
      const float dy = -__fdiv_rn(__fsub_rn((float)row + 0.5f, (float)p.H * 0.5f), p.focal);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        o[a] = __ldg(pose + 4 * a + 3);
        d[a] = fmaf(__ldg(pose + 4 * a + 2), -1.0f, fmaf(__ldg(pose + 4 * a + 1), dy, __ldg(pose + 4 * a) * dx));
      }
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        o[a] = __ldg(p.src_o + 3 * idx + a);
        d[a] = __ldg(p.src_d + 3 * idx + a);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p.rays_o[3 * i + a] = o[a];
      p.rays_d[3 * i + a] = d[a];
    }
  }
  if (p.pixels != nullptr && p.pixels_out != nullptr)
    for (int c = 0; c < p.C; ++c) p.pixels_out[i * p.C + c] = __ldg(p.pixels + idx * p.C + c);
}

}  // namespace

cudaError_t launch_sample_rays(long long n, long long k, int H, int W, int C, float focal, unsigned long long seed,
                               unsigned long long offset, const float* poses, const float* src_o, const float* src_d,
                               const float* pixels, const long long* idx_in, long long* idx_out, float* rays_o, float* rays_d,
                               float* pixels_out, cudaStream_t stream) {
  SamplerParams p{n, k, H, W, C, focal, seed, offset, poses, src_o, src_d, pixels, idx_in, idx_out, rays_o, rays_d, pixels_out};
  sample_rays_kernel<<<(unsigned)((k + 255) / 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace voxe
